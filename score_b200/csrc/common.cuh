// Shared device helpers for the SCoRe hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define FULL_MASK 0xffffffffu

namespace score {

// Step-scoped scalars the kernels read from device memory, so a captured CUDA graph can be
// replayed with new hyper-parameters by overwriting this one struct.
struct Hyper {
    float lr;
    float reg_lambda;
    float keep_prob;      // 1.0 disables dropout
    float alpha;          // TF Adam: lr * sqrt(1 - beta2^t) / (1 - beta1^t)
    float inv_batch;      // 1 / global batch (loss mean)
    uint32_t seed_lo, seed_hi;
    int32_t step;         // 1-based optimizer step this launch performs
    int32_t batch;        // B of this launch
    int32_t train;        // 1: training step (dropout active when keep_prob < 1)
    int32_t seq;          // launch sequence number (every upload): selects the ping-pong claim counter
    int32_t sample_base;  // global index of this launch's sample 0 (data-parallel: rank * per-rank batch): keys the dropout stream
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

// 16-byte async global->shared copy, L2 only (embedding rows are random: no L1 reuse).
// src_bytes == 0 zero-fills the destination (dummy node id 0 -> zero vector).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Programmatic dependent launch (sm_90+), an opt-in experiment (SCORE_PDL=1, kernels.h: launch_chain): the kernels of
// the step's critical chain can be launched with cudaLaunchAttributeProgrammaticStreamSerialization, so a kernel's CTAs
// are scheduled while its predecessor drains; pdl_enter() - the first statement of every such kernel - lets the
// successor be scheduled in turn and then waits until the predecessor grid has completed and its writes are visible.
// Launched without the attribute (the default) both instructions are no-ops.
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// Philox4x32-10 counter RNG (dropout masks, weight init): stateless, keyed by (seed, counter).
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
// uniform [0,1) for element `idx` of stream (`stream_id`, `step`)
__device__ __forceinline__ float philox_uniform(uint32_t seed_lo, uint32_t seed_hi, uint32_t stream_id,
                                                uint32_t step, uint64_t idx) {
    uint4 c = make_uint4((uint32_t)(idx >> 2), (uint32_t)(idx >> 34), stream_id, step);
    uint4 r = philox4x32(c, make_uint2(seed_lo, seed_hi));
    uint32_t w = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
    return (float)(w >> 8) * (1.0f / 16777216.0f);
}

// the four uniforms of elements 4*idx4 .. 4*idx4+3 of a stream: ONE Philox block (same values as philox_uniform)
__device__ __forceinline__ float4 philox_uniform4(uint32_t seed_lo, uint32_t seed_hi, uint32_t stream_id, uint32_t step,
                                                  uint64_t idx4) {
    uint4 c = make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), stream_id, step);
    uint4 r = philox4x32(c, make_uint2(seed_lo, seed_hi));
    const float s = 1.0f / 16777216.0f;
    return make_float4((float)(r.x >> 8) * s, (float)(r.y >> 8) * s, (float)(r.z >> 8) * s, (float)(r.w >> 8) * s);
}

}  // namespace score
