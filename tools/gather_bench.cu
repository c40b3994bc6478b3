// Microbenchmark of the select kernel's block gather in isolation: W warps per SM, every lane owns one "tree" (an arena of
// `cap` 32-byte slots) and repeatedly fetches a random 224-byte node block of it into shared memory, then spends `work`
// dependent FMAs on it (stand-in for the scoring).  Variants: 0 = cooperative cp.async (2 trees per instruction, as in
// k_select_f), 1 = per-lane 7 x LDG.256 into registers, 2 = per-lane 14 x cp.async 16 B (uncoalesced), 3 = cooperative cp.async
// with TWO blocks per lane in flight (software pipelining depth 2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather tools/gather_bench.cu && /tmp/gather
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31); }
constexpr int CTA = 64, ROW = 15;
template <int VAR>
__global__ void __launch_bounds__(CTA, 7) k_gather(const uint4 *pool, uint32_t cap, int levels, int work, uint64_t seed, float *sink) {
    __shared__ uint4 stage[2][CTA / 32][32][ROW];
    const int tid = blockIdx.x * CTA + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xFFFFFFFFu;
    const int part = lane & 15;
    const uint32_t env0 = tid - lane;
    const uint32_t tree_chunk0 = (env0 + (lane >> 4)) * cap * 2u + part, chunk_step = cap * 4u;
    float acc = 0.f;
    uint32_t off = (uint32_t)(mix(seed ^ tid) % (cap - 8));
    auto issue = [&](uint32_t o, int buf) {
        const uint32_t x = (o << 5) | 14u;
        uint32_t base = tree_chunk0;
        const unsigned sp = (unsigned)__cvta_generic_to_shared(&stage[buf][warp][lane >> 4][part]);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint32_t xt = __shfl_sync(FULL, x, 2 * i + (lane >> 4));
            if ((uint32_t)part < (xt & 15u)) {
                const uint4 *gp = pool + (base + (xt >> 4));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sp + (unsigned)(i * 2 * ROW * 16)), "l"(gp) : "memory");
            }
            base += chunk_step;
        }
    };
    if (VAR == 3) { issue(off, 0); asm volatile("cp.async.commit_group;" ::: "memory"); }
    for (int l = 0; l < levels; ++l) {
        uint4 v[7];
        const uint4 *mine = pool + ((size_t)tid * cap + off) * 2;
        if (VAR == 0) {
            issue(off, 0);
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 7; ++c) v[c] = stage[0][warp][lane][2 * c + 1];
        } else if (VAR == 1) {
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                uint32_t a, b, cc, d, e, f, g, h;
                asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(a), "=r"(b), "=r"(cc), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(mine + 2 * c));
                v[c] = make_uint4(e, f, g, h ^ a ^ b ^ cc ^ d);
            }
        } else if (VAR == 2) {
            const unsigned sp = (unsigned)__cvta_generic_to_shared(&stage[0][warp][lane][0]);
#pragma unroll
            for (int c = 0; c < 14; ++c) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sp + c * 16), "l"(mine + c) : "memory");
            asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 7; ++c) v[c] = stage[0][warp][lane][2 * c + 1];
        } else {
            // depth-2 pipeline: the NEXT block (address known in advance here - an upper bound for what prefetching could give)
            const uint32_t noff = (uint32_t)(mix(seed ^ tid ^ ((uint64_t)(l + 1) << 32)) % (cap - 8));
            issue(noff, (l + 1) & 1);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 7; ++c) v[c] = stage[l & 1][warp][lane][2 * c + 1];
            off = noff;
        }
        float x = __uint_as_float(v[0].x & 0x3FFFFFFF) + 1.0f;
#pragma unroll
        for (int c = 1; c < 7; ++c) x += __uint_as_float(v[c].y & 0x3FFFFFFF);
        for (int i = 0; i < work; ++i) x = fmaf(x, 1.0000001f, 0.5f);       // dependent chain
        acc += x;
        if (VAR != 3) off = (uint32_t)(mix(seed ^ tid ^ ((uint64_t)(l + 1) << 32) ^ (uint64_t)(__float_as_uint(x) & 1)) % (cap - 8));   // depends on the data
        __syncwarp();
    }
    if (acc == 123.456f) sink[0] = acc;
}
template <int VAR> void run(const uint4 *pool, uint32_t cap, int trees, int work, float *sink, const char *name) {
    const int levels = 48;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_gather<VAR><<<trees / CTA, CTA>>>(pool, cap, 4, work, 1, sink);
    cudaEventRecord(e0);
    k_gather<VAR><<<trees / CTA, CTA>>>(pool, cap, levels, work, 2, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)trees * levels * 224;
    printf("%-44s trees %6d work %4d: %7.2f us/level  %7.1f GB/s\n", name, trees, work, ms * 1e3 / levels, bytes / ms / 1e6);
}
int main() {
    const uint32_t cap = 2048; const int max_trees = 262144;
    uint4 *pool; float *sink; cudaMalloc(&pool, (size_t)max_trees * cap * 32); cudaMalloc(&sink, 64); cudaMemset(pool, 1, (size_t)max_trees * cap * 32);
    for (int trees : {65536, 131072, 262144})
        for (int work : {0, 256, 1024}) {
            run<0>(pool, cap, trees, work, sink, "cooperative cp.async (k_select_f)");
            run<1>(pool, cap, trees, work, sink, "per-lane 7 x LDG.256");
            run<2>(pool, cap, trees, work, sink, "per-lane 14 x cp.async.16");
            run<3>(pool, cap, trees, work, sink, "cooperative cp.async, next block prefetched");
        }
    return 0;
}
