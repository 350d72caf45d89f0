// 2-hop graph construction on the GPU: replaces GraphStore.construct_coll_2hop
// (code/graph_storage.py:127-246), which walks every (node, slice) in Python over MongoDB documents.
//
// Reference semantics, per node and slice t >= start_time (items first, then users - the order matters):
//   own = node['1hop'][t];  if len(own) > max_1hop: random.shuffle(own) IN PLACE (the shuffled list is what the new
//   document stores as '1hop', and what later look-ups of this node see), keep the first max_1hop;
//   for every kept neighbor n: deg = len(n['1hop'][t]);  deg <= 1: skipped;  deg <= max_1hop: the whole list;
//   else its first max_1hop entries;  'degrees' repeats deg once per appended id;
//   if the concatenation is longer than max_2hop: a random permutation (np.random.choice without replacement), the
//   first max_2hop entries of the permuted ids / degrees are kept.
// Item documents are built first and read the users' lists in their ORIGINAL order; user documents are built afterwards
// and read the items' lists AFTER the item phase shuffled them.
//
// Randomness is injected the way the sampler's is (sampler.cu): the reference uses Python's and NumPy's global
// generators, so its permutations cannot be reproduced; here the permutation of the k-th shuffle (k-th choice) call -
// calls counted in the reference's processing order - is the stable ascending argsort of the Philox4x32-10 uniforms
// u_j = philox_uniform(seed, stream 11 (12), k, j): new[r] = old[perm[r]].  tests/golden/hop2_reference.npz holds the
// output of the reference's own method run with random.shuffle / np.random.choice fed these permutations.
#include "../../include/score_b200.h"
#include "kernels.h"

#include <string>

namespace score {
namespace {

constexpr uint32_t STREAM_SHUFFLE = 11u, STREAM_CHOICE = 12u;

// ---- exclusive scan of n int32 values into OutT (three launches; block = 1024 values)
template <typename OutT>
__global__ void __launch_bounds__(1024) scan_block_kernel(const int32_t* __restrict__ in, int64_t n, OutT* __restrict__ out,
                                                           OutT* __restrict__ block_sums) {
    __shared__ OutT wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const OutT v = i < n ? (OutT)in[i] : (OutT)0;
    OutT x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const OutT y = __shfl_up_sync(FULL_MASK, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        OutT w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const OutT y = __shfl_up_sync(FULL_MASK, w, o);
            if (lane >= o) w += y;
        }
        wsum[lane] = w;
    }
    __syncthreads();
    if (i < n) out[i] = (warp ? wsum[warp - 1] : (OutT)0) + (x - v);
    if (threadIdx.x == 1023) block_sums[blockIdx.x] = wsum[31];
}
// one CTA: exclusive scan of the block sums in place, total -> block_sums[nblocks]
template <typename OutT>
__global__ void __launch_bounds__(1024) scan_sums_kernel(OutT* __restrict__ block_sums, int64_t nblocks) {
    __shared__ OutT wsum[32];
    __shared__ OutT carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < nblocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const OutT v = i < nblocks ? block_sums[i] : (OutT)0;
        OutT x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const OutT y = __shfl_up_sync(FULL_MASK, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            OutT w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const OutT y = __shfl_up_sync(FULL_MASK, w, o);
                if (lane >= o) w += y;
            }
            wsum[lane] = w;
        }
        __syncthreads();
        const OutT carry = carry_s;
        if (i < nblocks) block_sums[i] = carry + (warp ? wsum[warp - 1] : (OutT)0) + (x - v);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[nblocks] = carry_s;
}
template <typename OutT>
__global__ void __launch_bounds__(1024) scan_add_kernel(OutT* __restrict__ out, int64_t n, const OutT* __restrict__ block_sums) {
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    if (i < n) out[i] += block_sums[blockIdx.x];
    if (i == n - 1 || (n == 0 && i == 0)) out[n] = block_sums[gridDim.x];   // total at out[n]
}
// out[0..n] = exclusive scan of in[0..n) (out[n] = total); sums: (n + 1023) / 1024 + 1 elements of scratch
template <typename OutT>
void exclusive_scan(cudaStream_t st, const int32_t* in, int64_t n, OutT* out, OutT* sums) {
    const int64_t nb = (n + 1023) / 1024;
    if (nb == 0) { cudaMemsetAsync(out, 0, sizeof(OutT), st); return; }
    scan_block_kernel<OutT><<<(unsigned)nb, 1024, 0, st>>>(in, n, out, sums);
    scan_sums_kernel<OutT><<<1, 1024, 0, st>>>(sums, nb);
    scan_add_kernel<OutT><<<(unsigned)nb, 1024, 0, st>>>(out, n, sums);
}

struct H2 {
    int n_user, n_item, S, start, max1, max2;
    int64_t n_lists;            // (n_user + n_item + 1) * S, list index = node * S + t
    const int64_t* off1;
    uint32_t seed_lo, seed_hi;
};
__device__ __forceinline__ bool list_is_item(const H2& g, int64_t idx) { return idx >= (int64_t)(g.n_user + 1) * g.S; }

// lists the reference shuffles: slice >= start_time, longer than max_1hop
__global__ void hop2_flag_long_kernel(H2 g, int32_t* __restrict__ flag) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.n_lists) return;
    const int t = (int)(idx % g.S);
    const int64_t len = g.off1[idx + 1] - g.off1[idx];
    flag[idx] = (idx >= g.S && t >= g.start && len > g.max1) ? 1 : 0;
}
__global__ void hop2_compact_kernel(int64_t n, const int32_t* __restrict__ flag, const int32_t* __restrict__ excl, int64_t* __restrict__ list) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < n && flag[idx]) list[excl[idx]] = idx;
}
// call number of the k-th flagged list in the reference's processing order (items first, then users); CSR order is
// users first: rank_csr counts flagged lists before this one, n_user_flagged = flagged user lists
__device__ __forceinline__ uint32_t call_number(bool is_item, int rank_csr, int n_user_flagged, int n_flagged) {
    return is_item ? (uint32_t)(rank_csr - n_user_flagged) : (uint32_t)(rank_csr + (n_flagged - n_user_flagged));
}

// one CTA per shuffled list: new[rank_j] = old[j], rank_j = #{i : u_i < u_j or (u_i == u_j and i < j)}
constexpr int SHUF_THREADS = 256, SHUF_PER_THREAD = 8, SHUF_CHUNK = SHUF_THREADS * SHUF_PER_THREAD;
__global__ void __launch_bounds__(SHUF_THREADS) hop2_shuffle_kernel(H2 g, const int64_t* __restrict__ long_lists, const int32_t* __restrict__ excl,
                                                                    const int32_t* __restrict__ ids_old, int32_t* __restrict__ ids_new) {
    __shared__ float tile[SHUF_CHUNK];
    const int64_t idx = long_lists[blockIdx.x];
    const int64_t o = g.off1[idx];
    const int n = (int)(g.off1[idx + 1] - o);
    const int n_user_flagged = excl[(int64_t)(g.n_user + 1) * g.S], n_flagged = excl[g.n_lists];
    const uint32_t k = call_number(list_is_item(g, idx), blockIdx.x, n_user_flagged, n_flagged);
    for (int j0 = 0; j0 < n; j0 += SHUF_CHUNK) {
        float uj[SHUF_PER_THREAD]; int rank[SHUF_PER_THREAD];
#pragma unroll
        for (int q = 0; q < SHUF_PER_THREAD; ++q) {
            const int j = j0 + q * SHUF_THREADS + threadIdx.x;
            uj[q] = j < n ? philox_uniform(g.seed_lo, g.seed_hi, STREAM_SHUFFLE, k, (uint64_t)j) : 2.f;
            rank[q] = 0;
        }
        for (int i0 = 0; i0 < n; i0 += SHUF_CHUNK) {
            __syncthreads();
            for (int q = threadIdx.x; q < SHUF_CHUNK; q += SHUF_THREADS)
                tile[q] = i0 + q < n ? philox_uniform(g.seed_lo, g.seed_hi, STREAM_SHUFFLE, k, (uint64_t)(i0 + q)) : 3.f;
            __syncthreads();
            const int lim = min(SHUF_CHUNK, n - i0);
            for (int i = 0; i < lim; ++i) {
                const float ui = tile[i];
#pragma unroll
                for (int q = 0; q < SHUF_PER_THREAD; ++q) {
                    const int j = j0 + q * SHUF_THREADS + threadIdx.x;
                    rank[q] += (ui < uj[q] || (ui == uj[q] && i0 + i < j)) ? 1 : 0;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < SHUF_PER_THREAD; ++q) {
            const int j = j0 + q * SHUF_THREADS + threadIdx.x;
            if (j < n) ids_new[o + rank[q]] = ids_old[o + j];
        }
    }
}

// neighbor list of node `nb` in slice t as the reference sees it at that point: the users' lists in their original
// order while the item documents are built, the items' lists after their shuffle while the user documents are built
__device__ __forceinline__ void nbr_list(const H2& g, bool own_is_item, int32_t nb, int t, const int32_t* ids_old, const int32_t* ids_new,
                                         const int32_t** lst, int* deg) {
    const int64_t j = (int64_t)nb * g.S + t;
    const int64_t o = g.off1[j];
    *deg = (int)(g.off1[j + 1] - o);
    *lst = (own_is_item ? ids_old : ids_new) + o;
}

__global__ void hop2_count_kernel(H2 g, const int32_t* __restrict__ ids_old, const int32_t* __restrict__ ids_new,
                                  int32_t* __restrict__ len2, int32_t* __restrict__ flag2) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.n_lists) return;
    const int t = (int)(idx % g.S);
    int total = 0;
    if (idx >= g.S && t >= g.start) {
        const int64_t o = g.off1[idx];
        const int m = (int)min((int64_t)g.max1, g.off1[idx + 1] - o);
        const bool is_item = list_is_item(g, idx);
        for (int q = 0; q < m; ++q) {
            const int32_t* lst; int deg;
            nbr_list(g, is_item, ids_new[o + q], t, ids_old, ids_new, &lst, &deg);
            total += deg <= 1 ? 0 : min(deg, g.max1);
        }
    }
    len2[idx] = min(total, g.max2);
    flag2[idx] = total > g.max2 ? 1 : 0;
}

// one warp per (node, slice): the concatenation in shared memory (<= max1 * max1 entries), permuted when it exceeds max_2hop
__global__ void __launch_bounds__(256) hop2_fill_kernel(H2 g, const int32_t* __restrict__ ids_old, const int32_t* __restrict__ ids_new,
                                                        const int64_t* __restrict__ off2, const int32_t* __restrict__ excl2,
                                                        int32_t* __restrict__ ids2, int32_t* __restrict__ deg2, int cap) {
    extern __shared__ int32_t smh[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t idx = (int64_t)blockIdx.x * 8 + warp;
    if (idx >= g.n_lists) return;
    const int64_t o2 = off2[idx];
    const int n_out = (int)(off2[idx + 1] - o2);
    if (n_out == 0) return;
    int32_t* cid = smh + warp * 2 * cap;
    int32_t* cdg = cid + cap;
    const int t = (int)(idx % g.S);
    const int64_t o = g.off1[idx];
    const int m = (int)min((int64_t)g.max1, g.off1[idx + 1] - o);
    const bool is_item = list_is_item(g, idx);
    int total = 0;
    for (int q = 0; q < m; ++q) {   // every lane walks the (short) own list; the copies are spread over the lanes
        const int32_t* lst; int deg;
        nbr_list(g, is_item, ids_new[o + q], t, ids_old, ids_new, &lst, &deg);
        const int c = deg <= 1 ? 0 : min(deg, g.max1);
        for (int e = lane; e < c; e += 32) { cid[total + e] = lst[e]; cdg[total + e] = deg; }
        total += c;
    }
    __syncwarp();
    if (total <= g.max2) {
        for (int e = lane; e < total; e += 32) { ids2[o2 + e] = cid[e]; deg2[o2 + e] = cdg[e]; }
        return;
    }
    // permuted[r] = concat[perm[r]], r < max_2hop; perm = stable argsort of the call's uniforms
    const int n_user_flagged = excl2[(int64_t)(g.n_user + 1) * g.S], n_flagged = excl2[g.n_lists];
    const uint32_t k = call_number(is_item, excl2[idx], n_user_flagged, n_flagged);
    for (int j = lane; j < total; j += 32) {
        const float uj = philox_uniform(g.seed_lo, g.seed_hi, STREAM_CHOICE, k, (uint64_t)j);
        int rank = 0;
        for (int i = 0; i < total; ++i) {
            const float ui = philox_uniform(g.seed_lo, g.seed_hi, STREAM_CHOICE, k, (uint64_t)i);
            rank += (ui < uj || (ui == uj && i < j)) ? 1 : 0;
        }
        if (rank < g.max2) { ids2[o2 + rank] = cid[j]; deg2[o2 + rank] = cdg[j]; }
    }
}

thread_local std::string g_hop2_error;

#define HCK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_hop2_error = std::string(#call) + " failed: " + cudaGetErrorString(e_);              \
            rc = SCORE_ERR_CUDA;                                                                   \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)

}  // namespace
}  // namespace score

using namespace score;

extern "C" {

const char* score_graph_build_2hop_error(void) { return g_hop2_error.c_str(); }

int score_graph_build_2hop(const ScoreHop2Desc* d, int device, int32_t* hop1_ids_out, int64_t* hop2_off_out,
                           int32_t* hop2_ids_out, int32_t* hop2_deg_out, int64_t capacity, int64_t* n2_out) {
    if (!d || !d->hop1_off || !hop2_off_out || !n2_out || d->n_user < 0 || d->n_item < 0 || d->n_slices < 1 ||
        d->start_time < 0 || d->max_1hop < 1 || d->max_1hop > 32 || d->max_2hop < 1) {
        g_hop2_error = "score_graph_build_2hop: bad argument (max_1hop must be 1..32)";
        return SCORE_ERR_ARG;
    }
    int rc = SCORE_OK;
    H2 g{};
    g.n_user = d->n_user; g.n_item = d->n_item; g.S = d->n_slices; g.start = d->start_time; g.max1 = d->max_1hop; g.max2 = d->max_2hop;
    g.n_lists = (int64_t)(d->n_user + d->n_item + 1) * d->n_slices;
    g.seed_lo = (uint32_t)d->seed; g.seed_hi = (uint32_t)(d->seed >> 32);
    const int64_t L = g.n_lists, n1 = d->hop1_off[L];
    if (n1 > 0 && !d->hop1_ids) { g_hop2_error = "score_graph_build_2hop: hop1_ids is NULL"; return SCORE_ERR_ARG; }
    // the kernels index the offset array with the neighbor ids: validate the CSR arrays here, on the host, where they live
    // (the reference fails with an IndexError on a neighbor that has no document, graph_storage.py:160,208)
    if (d->hop1_off[0] != 0) { g_hop2_error = "score_graph_build_2hop: hop1_off[0] must be 0"; return SCORE_ERR_ARG; }
    for (int64_t i = 0; i < L; ++i)
        if (d->hop1_off[i + 1] < d->hop1_off[i]) { g_hop2_error = "score_graph_build_2hop: hop1_off is not ascending"; return SCORE_ERR_ARG; }
    {
        const int32_t hi = (int32_t)(d->n_user + d->n_item);
        for (int64_t i = 0; i < n1; ++i)
            if (d->hop1_ids[i] < 0 || d->hop1_ids[i] > hi) {
                g_hop2_error = "score_graph_build_2hop: neighbor id " + std::to_string(d->hop1_ids[i]) + " at position " + std::to_string(i) +
                               " is outside 0.." + std::to_string(hi);
                return SCORE_ERR_ID_RANGE;
            }
    }
    const int64_t nb = (L + 1023) / 1024 + 1;
    int64_t *off1 = nullptr, *off2 = nullptr, *sums64 = nullptr, *long_lists = nullptr;
    int32_t *ids_old = nullptr, *ids_new = nullptr, *flag = nullptr, *excl = nullptr, *sums32 = nullptr, *len2 = nullptr, *flag2 = nullptr,
            *excl2 = nullptr, *ids2 = nullptr, *deg2 = nullptr;
    cudaStream_t st = nullptr;
    int n_long = 0;
    int64_t n2 = 0;
    const unsigned grid = (unsigned)((L + 255) / 256);
    const int cap = d->max_1hop * d->max_1hop;
    HCK(cudaSetDevice(device));
    HCK(cudaStreamCreate(&st));
    HCK(cudaMalloc(&off1, sizeof(int64_t) * (L + 1)));
    HCK(cudaMalloc(&ids_old, sizeof(int32_t) * (n1 + 1)));
    HCK(cudaMalloc(&ids_new, sizeof(int32_t) * (n1 + 1)));
    HCK(cudaMalloc(&flag, sizeof(int32_t) * L)); HCK(cudaMalloc(&excl, sizeof(int32_t) * (L + 1))); HCK(cudaMalloc(&sums32, sizeof(int32_t) * nb));
    HCK(cudaMalloc(&len2, sizeof(int32_t) * L)); HCK(cudaMalloc(&flag2, sizeof(int32_t) * L)); HCK(cudaMalloc(&excl2, sizeof(int32_t) * (L + 1)));
    HCK(cudaMalloc(&off2, sizeof(int64_t) * (L + 1))); HCK(cudaMalloc(&sums64, sizeof(int64_t) * nb));
    HCK(cudaMemcpyAsync(off1, d->hop1_off, sizeof(int64_t) * (L + 1), cudaMemcpyHostToDevice, st));
    if (n1 > 0) HCK(cudaMemcpyAsync(ids_old, d->hop1_ids, sizeof(int32_t) * n1, cudaMemcpyHostToDevice, st));
    if (n1 > 0) HCK(cudaMemcpyAsync(ids_new, ids_old, sizeof(int32_t) * n1, cudaMemcpyDeviceToDevice, st));
    g.off1 = off1;
    // (1) the in-place shuffles of the long lists
    hop2_flag_long_kernel<<<grid, 256, 0, st>>>(g, flag);
    exclusive_scan<int32_t>(st, flag, L, excl, sums32);
    HCK(cudaMemcpyAsync(&n_long, excl + L, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    HCK(cudaStreamSynchronize(st));
    if (n_long > 0) {
        HCK(cudaMalloc(&long_lists, sizeof(int64_t) * n_long));
        hop2_compact_kernel<<<grid, 256, 0, st>>>(L, flag, excl, long_lists);
        hop2_shuffle_kernel<<<(unsigned)n_long, SHUF_THREADS, 0, st>>>(g, long_lists, excl, ids_old, ids_new);
    }
    // (2) sizes of the 2-hop lists, (3) their content
    hop2_count_kernel<<<grid, 256, 0, st>>>(g, ids_old, ids_new, len2, flag2);
    exclusive_scan<int64_t>(st, len2, L, off2, sums64);
    exclusive_scan<int32_t>(st, flag2, L, excl2, sums32);
    HCK(cudaMemcpyAsync(&n2, off2 + L, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    HCK(cudaMemcpyAsync(hop2_off_out, off2, sizeof(int64_t) * (L + 1), cudaMemcpyDeviceToHost, st));
    HCK(cudaStreamSynchronize(st));
    *n2_out = n2;
    if (hop1_ids_out && n1 > 0) HCK(cudaMemcpyAsync(hop1_ids_out, ids_new, sizeof(int32_t) * n1, cudaMemcpyDeviceToHost, st));
    if (hop2_ids_out && hop2_deg_out) {
        if (capacity < n2) { g_hop2_error = "score_graph_build_2hop: output capacity too small (see *n2_out)"; rc = SCORE_ERR_ARG; goto done; }
        if (n2 > 0) {
            HCK(cudaMalloc(&ids2, sizeof(int32_t) * n2)); HCK(cudaMalloc(&deg2, sizeof(int32_t) * n2));
            const size_t smem = (size_t)8 * 2 * cap * sizeof(int32_t);
            if (smem > 48 * 1024) HCK(cudaFuncSetAttribute(hop2_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            hop2_fill_kernel<<<(unsigned)((L + 7) / 8), 256, smem, st>>>(g, ids_old, ids_new, off2, excl2, ids2, deg2, cap);
            HCK(cudaMemcpyAsync(hop2_ids_out, ids2, sizeof(int32_t) * n2, cudaMemcpyDeviceToHost, st));
            HCK(cudaMemcpyAsync(hop2_deg_out, deg2, sizeof(int32_t) * n2, cudaMemcpyDeviceToHost, st));
        }
    }
    HCK(cudaStreamSynchronize(st));
    HCK(cudaGetLastError());
done:
    for (void* p : {(void*)off1, (void*)off2, (void*)sums64, (void*)long_lists, (void*)ids_old, (void*)ids_new, (void*)flag, (void*)excl,
                    (void*)sums32, (void*)len2, (void*)flag2, (void*)excl2, (void*)ids2, (void*)deg2})
        if (p) cudaFree(p);
    if (st) cudaStreamDestroy(st);
    return rc;
}

}  // extern "C"
