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

// ---- fused exchange over peer memory (NVLink): instead of gathering into a send buffer and calling an all-to-all, the
// kernels below store every row straight into its destination on the peer GPU.  cm = the all-gathered count matrix
// [world][world + 1] (cm[r][o] = positions of rank r owned by rank o), resident on every device; all offsets derive from it:
//   send_off(r, o) = sum_{o' < o} cm[r][o']   first slot of owner o in rank r's send order (= its staged-table rows - 1)
//   recv_off(o, r) = sum_{r' < r} cm[r'][o]   first element of requester r in owner o's served list / gradient buffer
__device__ __forceinline__ void shard_offsets(const int32_t* __restrict__ cm, int world, int me, int* send_off_me, int* recv_off_me,
                                              int* send_off_peer_me) {
    // shared arrays of world + 1 ints each, filled by the first warps; send_off_peer_me[r] = send_off(r, me)
    for (int i = threadIdx.x; i <= world; i += blockDim.x) {
        int s = 0, rcv = 0;
        for (int j = 0; j < i; ++j) { s += cm[me * (world + 1) + j]; rcv += cm[j * (world + 1) + me]; }
        send_off_me[i] = s; recv_off_me[i] = rcv;
        if (i < world) {
            int sp = 0;
            for (int j = 0; j < me; ++j) sp += cm[i * (world + 1) + j];
            send_off_peer_me[i] = sp;
        }
    }
    __syncthreads();
}

// owner side: served element e (local row want[e], requested by rank r) -> row send_off(r, me) + k + 1 of r's staged table
__global__ void __launch_bounds__(256) shard_serve_push_kernel(const float* __restrict__ table, int es, int d, int64_t V,
                                                               const int32_t* __restrict__ want, const int32_t* __restrict__ cm,
                                                               int world, int me, ShardPeers peers, int32_t* __restrict__ err_flag) {
    __shared__ int send_off_me[SHARD_MAX_WORLD + 1], recv_off_me[SHARD_MAX_WORLD + 1], send_off_peer_me[SHARD_MAX_WORLD];
    shard_offsets(cm, world, me, send_off_me, recv_off_me, send_off_peer_me);
    const int lpr = d >> 2;
    const int64_t total = (int64_t)recv_off_me[world] * lpr;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(t / lpr), c = (int)(t - (int64_t)e * lpr);
        int r = 0;
        while (r + 1 < world && recv_off_me[r + 1] <= e) ++r;      // world <= 64: a short scan over shared memory
        int32_t id = want[e];
        if (id < 0 || id >= V) { atomicExch(err_flag, 1); id = 0; }
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (id != 0) v = *reinterpret_cast<const float4*>(table + (int64_t)id * es + c * 4);
        const int64_t slot = (int64_t)send_off_peer_me[r] + (e - recv_off_me[r]) + 1;
        *reinterpret_cast<float4*>(peers.p[r] + slot * d + c * 4) = v;
    }
}

// requester side: send slot s (position sel[s], owner o) -> element recv_off(o, me) + k of o's gradient buffer
__global__ void __launch_bounds__(256) shard_grad_push_kernel(const float* __restrict__ grad_rows, const int32_t* __restrict__ sel,
                                                              int d, const int32_t* __restrict__ cm, int world, int me, ShardPeers peers) {
    __shared__ int send_off_me[SHARD_MAX_WORLD + 1], recv_off_me[SHARD_MAX_WORLD + 1], send_off_peer_me[SHARD_MAX_WORLD];
    __shared__ int recv_off_at[SHARD_MAX_WORLD];      // recv_off(o, me) for every owner o
    shard_offsets(cm, world, me, send_off_me, recv_off_me, send_off_peer_me);
    for (int o = threadIdx.x; o < world; o += blockDim.x) {
        int rcv = 0;
        for (int j = 0; j < me; ++j) rcv += cm[j * (world + 1) + o];
        recv_off_at[o] = rcv;
    }
    __syncthreads();
    const int lpr = d >> 2;
    const int64_t total = (int64_t)send_off_me[world] * lpr;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(t / lpr), c = (int)(t - (int64_t)s * lpr);
        int o = 0;
        while (o + 1 < world && send_off_me[o + 1] <= s) ++o;
        const float4 v = *reinterpret_cast<const float4*>(grad_rows + (int64_t)sel[s] * d + c * 4);
        const int64_t dst = (int64_t)recv_off_at[o] + (s - send_off_me[o]);
        *reinterpret_cast<float4*>(peers.p[o] + dst * d + c * 4) = v;
    }
}

void launch_shard_serve_push(cudaStream_t st, const float* table, int es, int d, int64_t V, const int32_t* want, int64_t n_recv,
                             const int32_t* cm, int world, int me, const ShardPeers& peers, int32_t* err_flag) {
    int64_t want_ctas = (n_recv * (d >> 2) + 255) / 256;
    if (want_ctas > 148 * 16) want_ctas = 148 * 16;
    if (want_ctas < 1) want_ctas = 1;
    shard_serve_push_kernel<<<(unsigned)want_ctas, 256, 0, st>>>(table, es, d, V, want, cm, world, me, peers, err_flag);
    ++g_launch_count;
}
void launch_shard_grad_push(cudaStream_t st, const float* grad_rows, const int32_t* sel, int64_t n, int d, const int32_t* cm,
                            int world, int me, const ShardPeers& peers) {
    int64_t want_ctas = (n * (d >> 2) + 255) / 256;
    if (want_ctas > 148 * 16) want_ctas = 148 * 16;
    if (want_ctas < 1) want_ctas = 1;
    shard_grad_push_kernel<<<(unsigned)want_ctas, 256, 0, st>>>(grad_rows, sel, d, cm, world, me, peers);
    ++g_launch_count;
}

}  // namespace score
