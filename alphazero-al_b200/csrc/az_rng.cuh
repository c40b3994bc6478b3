// Counter-based RNG.  Everything random in the engine (leaf symmetry ids, rollouts, Dirichlet noise) is a pure
// function of (seed, epoch, stream, env, counter), so results do not depend on how games are sharded over
// GPUs/threads.  The reference uses a thread_local std::mt19937 per OpenMP thread (src/cpp/MCTS.h:13-17,
// BatchedMCTS.h:68-84); that stream cannot be reproduced, so the integer draws here are
// matched bit-for-bit by the C restatement (oracle/az_oracle.c: orc_rand) instead.
#pragma once
#include <stdint.h>
#include "az_games.cuh"

namespace az {

enum { STREAM_SYM = 0, STREAM_ROLLOUT = 1, STREAM_NOISE = 2 };

AZ_HD uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
AZ_HD uint64_t az_rand(uint64_t seed, uint64_t epoch, uint64_t stream, uint64_t env, uint64_t ctr) {
    uint64_t h = splitmix64(seed ^ (stream * 0xD6E8FEB86659FD93ULL));
    h = splitmix64(h ^ epoch);
    h = splitmix64(h ^ (env << 24) ^ ctr);
    return h;
}
// config-2 rollout policy hash; twin of orc_rollout_hash
AZ_HD uint64_t rollout_hash(uint64_t seed, uint64_t gidx, uint64_t ply) {
    return splitmix64(splitmix64(seed ^ 0xA5A5A5A55A5A5A5AULL) ^ (gidx << 8) ^ ply);
}

#if defined(__CUDACC__)
// ---- programmatic dependent launch (sm_90+) ------------------------------------------------------------------------
// The playout loop is a chain of dependent kernels (select -> evaluate -> back-prop, 51 times per move and shard): with the
// programmatic-stream-serialization launch attribute a kernel's CTAs may become resident while the tail of its predecessor is
// still running; it executes whatever does not depend on the predecessor (index arithmetic, staging the {log, sqrt} table) and
// then blocks in pdl_wait() until the predecessor grid has completed and its writes are visible.  Without the attribute (or
// without a kernel predecessor) both calls are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// ---- L2 residency hints --------------------------------------------------------------------------------------------
// The leaf records, az_leaf rows, policy rows and value rows are producer -> consumer buffers rewritten in place every
// iteration (select -> evaluator -> back-prop), ~37 MB at 65 536 trees x K = 4.  Accessed with the evict_last policy they
// stay resident in the 126 MB L2, so neither their write-back nor their re-read reaches HBM.
__device__ __forceinline__ uint64_t l2_keep_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_f32_keep(float *p, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ float ld_f32_keep(const float *p, uint64_t pol) {
    float v;
    asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint4 ld_u4_keep(const void *p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_words256_keep(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t dd, uint32_t e, uint32_t f, uint32_t g,
                                                 uint32_t h, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8}, %9;" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(dd), "r"(e), "r"(f),
                 "r"(g), "r"(h), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async16_keep(void *smem_dst, const void *gsrc, uint64_t pol) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sa), "l"(gsrc), "l"(pol) : "memory");
}
#endif

}  // namespace az
