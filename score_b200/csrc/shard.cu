// Row-sharded embedding table (BASELINE.json config 5): device side of the exchange plan.
//
// owner(id) = id % world, owner-local row = id / world + 1 (local row 0 is the dummy).  A rank's positions are grouped
// by owner - position order inside a group, so the order in which an owner receives and sums gradient rows is fixed -
// with one stable counting pass (the radix-sort pass of scatter.cu on the owner number).  Nothing here needs the host:
// the per-owner counts stay on the device and are exchanged by the caller (one all-gather of the count matrix).
#include "kernels.h"

namespace score {

constexpr int SHARD_MAX_WORLD = 64;

__global__ void __launch_bounds__(256) shard_owner_kernel(const int32_t* __restrict__ keys, int64_t n, int world,
                                                          int32_t* __restrict__ owner, int32_t* __restrict__ counts) {
    __shared__ int cnt[SHARD_MAX_WORLD + 1];
    for (int i = threadIdx.x; i <= world; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int32_t id = keys[i];
        const int own = id != 0 ? (int)((uint32_t)id % (uint32_t)world) : world;   // dummy positions: bucket `world`
        owner[i] = own;
        atomicAdd(&cnt[own], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= world; i += blockDim.x)
        if (cnt[i]) atomicAdd(&counts[i], cnt[i]);   // integer counts: the result does not depend on the order
}

// slot s of the grouped list holds position sel[s]; the first n_valid = n - counts[world] slots have a real row
__global__ void __launch_bounds__(256) shard_finish_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ sorted_pos,
                                                           int64_t n, const int32_t* __restrict__ counts, int world,
                                                           int32_t* __restrict__ send_rows, int32_t* __restrict__ sel,
                                                           int32_t* __restrict__ mini_keys) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int64_t nv = n - counts[world];
    const int32_t p = sorted_pos[s];
    sel[s] = p;
    if (s < nv) {
        send_rows[s] = (int32_t)((uint32_t)keys[p] / (uint32_t)world) + 1;
        mini_keys[p] = (int32_t)s + 1;     // row of the staged table: the returned rows arrive in send order
    } else {
        mini_keys[p] = 0;
    }
}

void launch_shard_plan(cudaStream_t st, SortBufs& sb, const int32_t* keys, int64_t n, int world, int32_t* owner,
                       int32_t* counts, int32_t* send_rows, int32_t* sel, int32_t* mini_keys) {
    cudaMemsetAsync(counts, 0, sizeof(int32_t) * (world + 1), st);
    const unsigned grid = (unsigned)((n + 255) / 256);
    shard_owner_kernel<<<grid, 256, 0, st>>>(keys, n, world, owner, counts);
    ++g_launch_count;
    const int out = launch_sort_passes(st, sb, owner, n, 0, 1);   // one stable 8-bit pass: owner <= 64
    shard_finish_kernel<<<grid, 256, 0, st>>>(keys, sb.vals[out], n, counts, world, send_rows, sel, mini_keys);
    ++g_launch_count;
}

// gradient rows in send order: out[s] = grad_rows[sel[s]] for the slots that have a real row (one group of d/4 lanes per slot)
__global__ void __launch_bounds__(256) shard_pack_grads_kernel(const float* __restrict__ grad_rows, const int32_t* __restrict__ sel,
                                                               const int32_t* __restrict__ counts, int world, int64_t n, int d,
                                                               float* __restrict__ out) {
    const int lpr = d >> 2;
    const int64_t nv = n - counts[world];
    const int64_t total = nv * lpr;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = t / lpr;
        const int c = (int)(t - s * lpr);
        const float4 v = *reinterpret_cast<const float4*>(grad_rows + (int64_t)sel[s] * d + c * 4);
        *reinterpret_cast<float4*>(out + s * d + c * 4) = v;
    }
}
void launch_shard_pack_grads(cudaStream_t st, const float* grad_rows, const int32_t* sel, const int32_t* counts, int world,
                             int64_t n, int d, float* out) {
    int64_t want = (n * (d >> 2) + 255) / 256;
    if (want > 148 * 16) want = 148 * 16;
    if (want < 1) want = 1;
    shard_pack_grads_kernel<<<(unsigned)want, 256, 0, st>>>(grad_rows, sel, counts, world, n, d, out);
    ++g_launch_count;
}

}  // namespace score
