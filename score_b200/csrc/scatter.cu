// Optimizer side of the SCoRe path: deterministic embedding-gradient scatter + Adam.
//
// The reference gets its embedding gradient by autodiff through six embedding_lookups on a
// dense [V,d] tensor (score.py:45-66) and applies tf.train.AdamOptimizer to it (score.py:98).
// Here the per-position gradient rows produced by the backward kernels are combined WITHOUT
// float atomics:  stable LSD radix sort of (row id, position)  ->  one thread group per unique
// row walks its run in position order (fixed summation order)  ->  fused TF-formulation Adam
// on that row's (var, m, v) with 128-bit accesses.
//
// Adam arithmetic uses explicit round-to-nearest intrinsics in the op order of TF's ApplyAdam
// kernel so that no FMA contraction changes a bit:
//     alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t)          (host, fp32)
//     m += (g - m) * (1 - beta1);  v += (g*g - v) * (1 - beta2);  var -= (m * alpha) / (sqrt(v) + eps)
#include <stdlib.h>

#include "kernels.h"

namespace score {

__device__ __forceinline__ void adam_elem(float& var, float& m, float& v, float g, float alpha) {
    const float omb1 = 1.0f - 0.9f, omb2 = 1.0f - 0.999f, eps = 1e-8f;
    m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), omb1));
    v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), omb2));
    var = __fsub_rn(var, __fdiv_rn(__fmul_rn(m, alpha), __fadd_rn(__fsqrt_rn(v), eps)));
}
__device__ __forceinline__ void adam4(float4& var, float4& m, float4& v, const float4& g, float alpha) {
    adam_elem(var.x, m.x, v.x, g.x, alpha);
    adam_elem(var.y, m.y, v.y, g.y, alpha);
    adam_elem(var.z, m.z, v.z, g.z, alpha);
    adam_elem(var.w, m.w, v.w, g.w, alpha);
}
__device__ __forceinline__ bool all_zero(const float4& a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f && a.w == 0.f; }

// ------------------------------------------------------------------------------------------ dense parameters
// 0.5 * sum v^2 over L2-regularised parameters: L2_PARTS blocks, fixed-shape tree per block; loss_final adds the
// per-block partials in index order (deterministic)
__global__ void l2_sum_kernel(const float* __restrict__ p, const uint8_t* __restrict__ flags, int n, float* out) {
    __shared__ float red[256];
    float s = 0.f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256)
        if (flags[i] & 1) s += p[i] * p[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0] * 0.5f;
}
void launch_l2_sum(cudaStream_t st, const float* params, const uint8_t* flags, int n, float* out) {
    l2_sum_kernel<<<L2_PARTS, 256, 0, st>>>(params, flags, n, out);
    ++g_launch_count;
}

// G[i] = fixed-order sum of the partial planes (+ derived range); with `ad` set the dense Adam step of element i follows
// in the same thread (training path: one launch instead of two on the critical path)
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int splits, int n, float* __restrict__ g,
                                       int64_t d_dst, int64_t d_a, int64_t d_b, int d_count, DenseAdamArgs ad) {
    pdl_enter();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && ad.p && ad.alpha_hist) ad.alpha_hist[ad.hp->step] = ad.hp->alpha;
    if (i >= n) return;
    // fixed-order sum over the planes, eight independent loads in flight
    auto plane_sum = [&](int64_t col) {
        float s = 0.f;
        int k = 0;
        for (; k + 8 <= splits; k += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = partials[(int64_t)(k + u) * n + col];
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        for (; k < splits; ++k) s += partials[(int64_t)k * n + col];
        return s;
    };
    float gi;
    if (d_dst >= 0 && i >= d_dst && i < d_dst + d_count) {
        // derived range: the same fixed-order sums its two source elements get, then their difference
        gi = plane_sum(d_a + (i - d_dst)) - plane_sum(d_b + (i - d_dst));
    } else {
        gi = plane_sum(i);
    }
    g[i] = gi;
    if (!ad.p) return;
    const uint8_t f = ad.flags[i];
    if (!(f & 2)) return;   // non-trainable (bn moving statistics)
    float var = ad.p[i], mm = ad.m[i], vv = ad.v[i];
    if (f & 1) gi = __fadd_rn(gi, __fmul_rn(ad.hp->reg_lambda, var));   // d/dv of reg_lambda * l2_loss(v)
    adam_elem(var, mm, vv, gi, ad.hp->alpha);
    ad.p[i] = var; ad.m[i] = mm; ad.v[i] = vv;
}
void launch_reduce_partials(cudaStream_t st, const float* partials, int splits, int n, float* g, int64_t derive_dst,
                            int64_t derive_a, int64_t derive_b, int derive_count, const DenseAdamArgs* adam) {
    DenseAdamArgs ad{};
    if (adam) ad = *adam;
    launch_chain(reduce_partials_kernel, dim3((n + 255) / 256), dim3(256), 0, st, partials, splits, n, g, derive_dst, derive_a, derive_b,
                                                            derive_count, ad);
    ++g_launch_count;
}

// inspection only (score_forward_backward): G += reg_lambda * p for L2-regularised parameters, so the exported
// gradient is the gradient of the full loss as tf.gradients would report it
__global__ void add_l2_kernel(float* __restrict__ g, const float* __restrict__ p, const uint8_t* __restrict__ flags, int n,
                              const Hyper* hp) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (flags[i] & 1)) g[i] = __fadd_rn(g[i], __fmul_rn(hp->reg_lambda, p[i]));
}
void launch_add_l2(cudaStream_t st, float* g, const float* p, const uint8_t* flags, int n, const Hyper* hp) {
    add_l2_kernel<<<(n + 255) / 256, 256, 0, st>>>(g, p, flags, n, hp);
    ++g_launch_count;
}

__global__ void dense_adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                  const float* __restrict__ g, const uint8_t* __restrict__ flags, int n,
                                  const Hyper* hp, float* alpha_hist) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && alpha_hist) alpha_hist[hp->step] = hp->alpha;
    if (i >= n) return;
    uint8_t f = flags[i];
    if (!(f & 2)) return;   // non-trainable (bn moving statistics)
    float var = p[i], mm = m[i], vv = v[i];
    float gi = g[i];
    if (f & 1) gi = __fadd_rn(gi, __fmul_rn(hp->reg_lambda, var));   // d/dv of reg_lambda * l2_loss(v)
    adam_elem(var, mm, vv, gi, hp->alpha);
    p[i] = var; m[i] = mm; v[i] = vv;
}
void launch_dense_adam(cudaStream_t st, float* p, float* m, float* v, const float* g, const uint8_t* flags,
                       int n, const Hyper* hp, float* alpha_hist) {
    dense_adam_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, m, v, g, flags, n, hp, alpha_hist);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------ radix sort
// Stable LSD radix sort of (key, value) pairs, 8-bit digits, three kernels per pass:
// per-tile digit histogram -> exclusive scan over (digit-major, tile-minor) -> ranked scatter.
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

size_t sort_hist_elems(int64_t n) {
    size_t total = (size_t)256 * ((n + SORT_TILE - 1) / SORT_TILE);
    return total + (total + 1023) / 1024 + 1;   // histogram + one total per 1024-entry scan chunk
}

// The tile kernels are grid-stride over the tiles so the grid can be capped (sort_grid_cap).
__global__ void sort_hist_kernel(const int32_t* __restrict__ keys, int64_t n, int shift, uint32_t* __restrict__ hist,
                                 int nblocks) {
    __shared__ uint32_t cnt[256];
    for (int tile = blockIdx.x; tile < nblocks; tile += gridDim.x) {
        cnt[threadIdx.x] = 0;
        __syncthreads();
        const int64_t base = (int64_t)tile * SORT_TILE;
#pragma unroll
        for (int r = 0; r < SORT_ITEMS; ++r) {
            int64_t i = base + r * SORT_THREADS + threadIdx.x;
            if (i < n) atomicAdd(&cnt[((uint32_t)keys[i] >> shift) & 255u], 1u);
        }
        __syncthreads();
        hist[(int64_t)threadIdx.x * nblocks + tile] = cnt[threadIdx.x];
    }
}

// Exclusive scan of the (digit-major, tile-minor) histogram in two fully parallel kernels:
//   sort_scan_local: each block of 1024 scans its own 1024-entry chunk in place and publishes the chunk total;
//   the consumer (sort_scatter_kernel) adds the prefix of the chunk totals on the fly (<= a few hundred adds).
__global__ void sort_scan_local_kernel(uint32_t* __restrict__ hist, int total, uint32_t* __restrict__ chunk_sums) {
    __shared__ uint32_t wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 1024 + threadIdx.x;
    const uint32_t v = (i < total) ? hist[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(FULL_MASK, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(FULL_MASK, w, o);
            if (lane >= o) w += y;
        }
        wsum[lane] = w;
    }
    __syncthreads();
    if (i < total) hist[i] = (warp ? wsum[warp - 1] : 0u) + (x - v);
    if (threadIdx.x == 1023) chunk_sums[blockIdx.x] = wsum[31];
}

// Ranked scatter of one tile.  The (key, value) pairs are first put in digit order INSIDE the tile (shared memory), then
// written out by consecutive threads: the pairs of one (tile, digit) run are consecutive in the output, so a warp's
// stores cover a few contiguous runs instead of 32 unrelated 4-byte sectors (the scattered version spent most of the
// CCMR-size sort - 9.8 M keys - on write sectors).
__global__ void __launch_bounds__(SORT_THREADS)
sort_scatter_kernel(const int32_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in,
                    int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out, int64_t n, int shift,
                    const uint32_t* __restrict__ hist, int nblocks, const uint32_t* __restrict__ chunk_sums) {
    constexpr int WARPS = SORT_THREADS / 32;
    __shared__ uint32_t whist[WARPS][256];
    __shared__ uint32_t gbase[256];      // global position of the tile's first pair of each digit, minus its local start
    __shared__ uint32_t lstart[256];     // start of each digit inside the tile
    __shared__ uint32_t wscan[WARPS];
    __shared__ int32_t skey[SORT_TILE], sval[SORT_TILE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int tile = blockIdx.x; tile < nblocks; tile += gridDim.x) {
    __syncthreads();   // the previous tile's readers of the shared arrays are done
    for (int i = threadIdx.x; i < WARPS * 256; i += SORT_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();
    const int64_t tbase = (int64_t)tile * SORT_TILE;
    const int64_t wbase = tbase + warp * (32 * SORT_ITEMS);
    const int tile_n = (int)min((int64_t)SORT_TILE, n - tbase);
    int32_t k[SORT_ITEMS];
    uint32_t lrank[SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        int64_t i = wbase + r * 32 + lane;
        bool ok = i < n;
        k[r] = ok ? keys_in[i] : 0;
        uint32_t digit = ok ? (((uint32_t)k[r] >> shift) & 255u) : (256u + lane);   // invalid lanes match only themselves
        uint32_t peers = __match_any_sync(FULL_MASK, digit);
        uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (ok && lane == leader) { old = whist[warp][digit]; whist[warp][digit] = old + __popc(peers); }
        old = __shfl_sync(FULL_MASK, old, leader);
        lrank[r] = old + rank;
        __syncwarp();
    }
    __syncthreads();
    {   // per digit: exclusive prefix over the warps, the tile's count, its start inside the tile (256-wide scan)
        const int dgt = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) { uint32_t c = whist[w][dgt]; whist[w][dgt] = run; run += c; }
        uint32_t incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) wscan[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) before += w < warp ? wscan[w] : 0u;
        const uint32_t ls = before + incl - run;
        lstart[dgt] = ls;
        const int64_t hidx = (int64_t)dgt * nblocks + tile;
        uint32_t pre = 0;
        const int chunk = (int)(hidx >> 10);
        for (int c = 0; c < chunk; ++c) pre += chunk_sums[c];   // prefix of the 1024-entry chunk totals
        gbase[dgt] = hist[hidx] + pre - ls;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        int64_t i = wbase + r * 32 + lane;
        if (i < n) {
            uint32_t digit = ((uint32_t)k[r] >> shift) & 255u;
            const uint32_t lp = lstart[digit] + whist[warp][digit] + lrank[r];
            skey[lp] = k[r];
            sval[lp] = vals_in ? vals_in[i] : (int32_t)i;
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < tile_n; e += SORT_THREADS) {
        const int32_t key = skey[e];
        const uint32_t pos = gbase[((uint32_t)key >> shift) & 255u] + (uint32_t)e;
        keys_out[pos] = key;
        vals_out[pos] = sval[e];
    }
    }   // tiles
}

// CTAs of the tile kernels: SCORE_SORT_CTAS caps them (profiling knob; measured on B200: capping the grid makes the
// gather a little faster but stretches the sort into the dense backward - the step is fastest uncapped).
static int sort_grid_cap() {
    static int cap = 0;
    if (!cap) {
        const char* e = getenv("SCORE_SORT_CTAS");
        cap = e ? atoi(e) : 0;
        if (cap <= 0) cap = 1 << 30;
    }
    return cap;
}
int sort_num_passes(int key_bits) {
    const int passes = (key_bits + 7) / 8;
    return passes < 1 ? 1 : passes;
}
// passes [p_begin, p_end) of the sort; pass 0 reads keys_in, pass p > 0 the output of pass p-1 (ping-pong p & 1)
int launch_sort_passes(cudaStream_t st, SortBufs& sb, const int32_t* keys_in, int64_t n, int p_begin, int p_end) {
    const int nblocks = (int)((n + SORT_TILE - 1) / SORT_TILE);
    const int grid = nblocks < sort_grid_cap() ? nblocks : sort_grid_cap();
    int out = (p_begin > 0) ? ((p_begin - 1) & 1) : 0;
    for (int p = p_begin; p < p_end; ++p) {
        const int32_t* kin = p == 0 ? keys_in : sb.keys[(p - 1) & 1];
        const int32_t* vin = p == 0 ? nullptr : sb.vals[(p - 1) & 1];
        out = p & 1;
        sort_hist_kernel<<<grid, SORT_THREADS, 0, st>>>(kin, n, 8 * p, sb.hist, nblocks);
        const int total = 256 * nblocks, nchunks = (total + 1023) / 1024;
        uint32_t* chunk_sums = sb.hist + total;   // tail of the histogram allocation
        sort_scan_local_kernel<<<nchunks, 1024, 0, st>>>(sb.hist, total, chunk_sums);
        sort_scatter_kernel<<<grid, SORT_THREADS, 0, st>>>(kin, vin, sb.keys[out], sb.vals[out], n, 8 * p, sb.hist,
                                                             nblocks, chunk_sums);
        g_launch_count += 3;
    }
    return out;
}
int launch_sort_pairs(cudaStream_t st, SortBufs& sb, const int32_t* keys_in, int64_t n, int key_bits) {
    return launch_sort_passes(st, sb, keys_in, n, 0, sort_num_passes(key_bits));
}

// LAZY mode: replay the zero-gradient steps (from, upto] of one row chunk, same op sequence as the sweep.
__device__ __forceinline__ void replay4(float4& var, float4& m, float4& v, int from, int upto,
                                        const float* __restrict__ alpha_hist) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = from + 1; s <= upto; ++s) adam4(var, m, v, z, alpha_hist[s]);
}
// ------------------------------------------------------------------------------------------ segment reduce + row Adam
// Real batches hold very hot rows next to cold ones: a side-feature id (category, gender, genre ...) occurs hundreds to
// thousands of times in one batch, a user or item id once or twice.  A serial walk of the hot runs would set the kernel
// time (ncu: with one thread group per run the Taobao-shape update took 60-75 us, all of it the 207-entry runs), so the
// runs of the sorted list are split into three tiers by length and every tier has its own fixed summation tree:
//   S  (<= 4 entries)     one group of d/4 lanes per run; the 32-byte descriptor holds the run's positions, so the
//                         optimizer state and all gradient rows are ONE batch of independent 16-byte loads;
//   M  (5 .. 32)          one team of 16 lanes (d <= 32) or one warp per run: group g of the team adds entries g, g+G,
//                         g+2G ... in ascending order, the G partial sums are added in group order (shuffles);
//   L  (> 32)             one CTA per run: the same with 256/LPR groups, the warps' sums added in warp order (smem).
// No float atomics anywhere; the tree of a run depends only on its length, so results are reproducible.
// emb_runs_kernel (on the sort stream, off the critical path) builds the three descriptor lists; the order inside a
// list comes from atomics and only decides which threads serve which run.
constexpr int RUN_S_MAX = 4;
// A run of more than RUN_CHUNK entries (a hot side-feature id, a Zipf head item: tens of thousands of positions) is cut
// into chunks of RUN_CHUNK entries, one CTA each; the CTA that finishes last adds the chunk sums IN CHUNK ORDER and
// applies the row's optimizer step.  The summation tree still depends on the run length only.
constexpr int RUN_CHUNK = 2048;
// Upper length of tier M.  A team of 16 lanes needs ceil(n/16) dependent load rounds for a run of n entries, a CTA one
// round for up to 256, so the break-even is around 32 entries.  Measured on B200 (Taobao shape, uniform ids, longest
// run 207): 32 and 512 give the same 30 us inside the step - the hot runs are not what bounds the kernel there - so
// the smaller value is kept for skewed id distributions, where the hot runs are longer.
// SCORE_RUN_M_MAX overrides (A/B knob); RUN_M_FLOOR sizes the descriptor lists for the smallest allowed value.
constexpr int RUN_M_DEFAULT = 32;
constexpr int RUN_M_FLOOR = 16;
static int run_m_max() {
    static int v = 0;
    if (!v) {
        const char* e = getenv("SCORE_RUN_M_MAX");
        v = e ? atoi(e) : RUN_M_DEFAULT;
        if (v < RUN_M_FLOOR) v = RUN_M_FLOOR;
    }
    return v;
}

__global__ void emb_runs_kernel(const int32_t* __restrict__ skeys, const int32_t* __restrict__ spos, int64_t n,
                                int4* __restrict__ runs, int4* __restrict__ runs_long, int64_t long_cap,
                                int32_t* __restrict__ counters, int m_max, int4* __restrict__ slotinfo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int32_t key = 0, prev = -1;
    if (i < n) { key = skeys[i]; prev = i > 0 ? skeys[i - 1] : -1; }
    const bool head = (i < n) && key != 0 && key != prev;
    const unsigned hmask = __ballot_sync(FULL_MASK, head);
    if (hmask == 0) return;
    if (lane == 0) atomicAdd(counters + 3, __popc(hmask));   // total number of runs (= unique rows)
    int32_t k[4] = {0, 0, 0, 0}, p[4] = {0, 0, 0, 0};
    if (head) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {   // independent loads
            k[u] = (i + 1 + u < n) ? skeys[i + 1 + u] : 0;
            p[u] = (i + u < n) ? spos[i + u] : 0;
        }
    }
    int n4 = 1;
#pragma unroll
    for (int u = 0; u < 3; ++u) n4 += (n4 == u + 1 && k[u] == key) ? 1 : 0;
    const bool more = head && n4 == 4 && k[3] == key;
    const bool is_s = head && !more;
    const unsigned smask = __ballot_sync(FULL_MASK, is_s);
    int base = 0;
    if (lane == 0 && smask) base = atomicAdd(counters, __popc(smask));
    base = __shfl_sync(FULL_MASK, base, 0);
    if (is_s) {
        const int64_t slot = base + __popc(smask & ((1u << lane) - 1u));
        runs[2 * slot] = make_int4(key, (int32_t)i, n4, p[0]);
        runs[2 * slot + 1] = make_int4(p[1], p[2], p[3], 0);
    }
    if (!more) return;
    // long run: entries i .. i+4 carry the key; gallop, then bisect, to the first index that does not
    int64_t lo = i + 4, stepw = 4;
    while (lo + stepw < n && skeys[lo + stepw] == key) { lo += stepw; stepw <<= 1; }
    int64_t hi = lo + stepw < n ? lo + stepw : n;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (skeys[mid] == key) lo = mid; else hi = mid;
    }
    const int32_t cnt = (int32_t)(hi - i);
    if (cnt <= m_max) {
        const int slot = atomicAdd(counters + 1, 1);
        runs_long[slot] = make_int4(key, (int32_t)i, cnt, 0);              // M list grows from the front
    } else if (cnt <= RUN_CHUNK || !slotinfo) {
        const int slot = atomicAdd(counters + 2, 1);
        runs_long[long_cap - 1 - slot] = make_int4(key, (int32_t)i, cnt, -1);   // L list grows from the back; -1: one chunk
    } else {
        const int nch = (cnt + RUN_CHUNK - 1) / RUN_CHUNK;
        const int slot = atomicAdd(counters + 2, nch), pb = atomicAdd(counters + 4, nch);
        for (int c = 0; c < nch; ++c) {
            const int len = min(RUN_CHUNK, cnt - c * RUN_CHUNK);
            runs_long[long_cap - 1 - (slot + c)] = make_int4(key, (int32_t)i + c * RUN_CHUNK, len, pb + c);
            slotinfo[pb + c] = make_int4(pb, nch, (int32_t)i, 0);
        }
    }
}
int64_t emb_runs_long_cap(int64_t n) { return n / (RUN_S_MAX + 1) + n / (RUN_M_FLOOR + 1) + n / RUN_CHUNK + 16; }
int64_t emb_runs_part_cap(int64_t n) { return 2 * (n / RUN_CHUNK) + 16; }   // sum over long runs of ceil(cnt / RUN_CHUNK) <= 2n / RUN_CHUNK
void launch_emb_runs(cudaStream_t st, const int32_t* skeys, const int32_t* spos, int64_t n, int32_t* runs,
                     int32_t* runs_long, int32_t* counters, int32_t* slotinfo) {
    cudaMemsetAsync(counters, 0, 8 * sizeof(int32_t), st);
    emb_runs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(skeys, spos, n, reinterpret_cast<int4*>(runs),
                                                                 reinterpret_cast<int4*>(runs_long), emb_runs_long_cap(n),
                                                                 counters, run_m_max(), reinterpret_cast<int4*>(slotinfo));
    ++g_launch_count;
}


static int num_sms_cached();
// ------------------------------------------------------------------------------------------ data-parallel exchange helpers
// Rank of every run head among the run heads of the sorted key list (exclusive count of heads before it): the slot of
// the run in a compact list that is still ascending by key.  Two kernels on the sort stream, off the critical path:
// per-tile head counts, then (prefix of the tile counts) + (rank inside the tile).
constexpr int HS_THREADS = 256;
constexpr int HS_ITEMS = 8;
constexpr int HS_TILE = HS_THREADS * HS_ITEMS;

__device__ __forceinline__ bool is_run_head(const int32_t* __restrict__ skeys, int64_t i, int64_t n) {
    if (i >= n) return false;
    const int32_t key = skeys[i];
    return key != 0 && (i == 0 || skeys[i - 1] != key);
}
__global__ void __launch_bounds__(HS_THREADS) head_count_kernel(const int32_t* __restrict__ skeys, int64_t n,
                                                                int32_t* __restrict__ tile_counts) {
    __shared__ int wcnt[HS_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * HS_TILE;
    int c = 0;
#pragma unroll
    for (int r = 0; r < HS_ITEMS; ++r) c += is_run_head(skeys, base + r * HS_THREADS + threadIdx.x, n) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
    if ((threadIdx.x & 31) == 0) wcnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < HS_THREADS / 32; ++w) t += wcnt[w];
        tile_counts[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(HS_THREADS) head_slot_kernel(const int32_t* __restrict__ skeys, int64_t n,
                                                               const int32_t* __restrict__ tile_counts,
                                                               int32_t* __restrict__ head_slot) {
    constexpr int WARPS = HS_THREADS / 32;
    __shared__ int red[WARPS];
    __shared__ int cnt[HS_ITEMS][WARPS];
    __shared__ int tile_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // heads in the tiles before this one
    int pre = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += HS_THREADS) pre += tile_counts[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(FULL_MASK, pre, o);
    if (lane == 0) red[warp] = pre;
    const int64_t base = (int64_t)blockIdx.x * HS_TILE;
    bool head[HS_ITEMS];
    unsigned bal[HS_ITEMS];
#pragma unroll
    for (int r = 0; r < HS_ITEMS; ++r) {
        head[r] = is_run_head(skeys, base + r * HS_THREADS + threadIdx.x, n);
        bal[r] = __ballot_sync(FULL_MASK, head[r]);
        if (lane == 0) cnt[r][warp] = __popc(bal[r]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) t += red[w];
        tile_base = t;
        int run = 0;   // exclusive prefix in (item, warp) order = ascending sorted index
        for (int r = 0; r < HS_ITEMS; ++r)
            for (int w = 0; w < WARPS; ++w) { const int c = cnt[r][w]; cnt[r][w] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < HS_ITEMS; ++r)
        if (head[r]) head_slot[base + r * HS_THREADS + threadIdx.x] = tile_base + cnt[r][warp] + __popc(bal[r] & ((1u << lane) - 1u));
}
int64_t head_slot_tiles(int64_t n) { return (n + HS_TILE - 1) / HS_TILE; }
void launch_head_slots(cudaStream_t st, const int32_t* skeys, int64_t n, int32_t* tile_counts, int32_t* head_slot) {
    const unsigned tiles = (unsigned)head_slot_tiles(n);
    if (!tiles) return;
    head_count_kernel<<<tiles, HS_THREADS, 0, st>>>(skeys, n, tile_counts);
    head_slot_kernel<<<tiles, HS_THREADS, 0, st>>>(skeys, n, tile_counts, head_slot);
    g_launch_count += 2;
}

// {number of unique rows, sequence number of the step} for the host, which sizes the exchange from it while the
// forward / backward kernels are still running
__global__ void dp_count_kernel(const int32_t* __restrict__ counters, const Hyper* hp, int32_t* __restrict__ slot) {
    slot[0] = counters[3];
    slot[1] = hp->seq;
}
void launch_dp_count(cudaStream_t st, const int32_t* counters, const Hyper* hp, int32_t* slot) {
    dp_count_kernel<<<1, 1, 0, st>>>(counters, hp, slot);
    ++g_launch_count;
}
// Everything of this rank's packed block except the exported rows, in one launch: header (unique-row count, step
// sequence number, loss total / L2 part), the dense gradient, zeros in the id slots past the count
__global__ void dp_pack_misc_kernel(int32_t* __restrict__ block, DpLayout L, const int32_t* __restrict__ counters,
                                    const Hyper* hp, const float* __restrict__ loss_dev, const float* __restrict__ g, int n_dense) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t cnt = counters[3];
    if (i == 0) {
        block[0] = cnt;
        block[1] = hp->seq;
        block[2] = __float_as_int(loss_dev[0]);
        block[3] = __float_as_int(loss_dev[1]);
    }
    if (i < n_dense) reinterpret_cast<float*>(block + L.dense_off)[i] = g[i];
    if (i >= cnt && i < L.cap) block[L.keys_off + i] = 0;
}
void launch_dp_pack_misc(cudaStream_t st, int32_t* block, const DpLayout& L, const int32_t* counters, const Hyper* hp,
                         const float* loss_dev, const float* g, int n_dense) {
    const int64_t n = L.cap > n_dense ? L.cap : n_dense;
    dp_pack_misc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(block, L, counters, hp, loss_dev, g, n_dense);
    ++g_launch_count;
}

// Peer-memory variant of the exchange (opt-in, SCORE_DP_P2P=1 in parallel.py): this rank's block is stored straight
// into every replica's gathered buffer over NVLink (16-byte stores, one grid-stride pass per destination) instead of
// going through an all-gather; a device-side barrier of the symmetric-memory handle orders it against the readers.
__global__ void __launch_bounds__(256) dp_push_kernel(const int4* __restrict__ src, int64_t n16, DpPeers peers) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const int4 v = src[i];
#pragma unroll 4
        for (int r = 0; r < peers.world; ++r) reinterpret_cast<int4*>(peers.dst[r])[i] = v;
    }
}
void launch_dp_push(cudaStream_t st, const int32_t* block, int64_t words, const DpPeers& peers) {
    const int64_t n16 = words / 4;   // a block is a multiple of 128 words
    int64_t want = (n16 + 255) / 256, cap = (int64_t)num_sms_cached() * 8;
    dp_push_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(reinterpret_cast<const int4*>(block), n16, peers);
    ++g_launch_count;
}

// Dense half of the data-parallel finish: G[i] = sum over ranks (rank order: every replica adds in the same order, so
// the replicas stay bit-identical whatever the collective does) of the gathered dense gradients, then TF-form Adam.
// Thread 0 also composes the global loss: sum of the ranks' data terms (each already scaled by 1/global_batch) + L2 once.
__global__ void dp_dense_adam_kernel(DpLayout L, float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                     float* __restrict__ g_out, const uint8_t* __restrict__ flags, int n, const Hyper* hp,
                                     float* alpha_hist, double* __restrict__ loss_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        if (alpha_hist) alpha_hist[hp->step] = hp->alpha;
        double s = 0.0;
        for (int r = 0; r < L.world; ++r) {
            const int32_t* blk = L.base + (int64_t)r * L.stride;
            s += (double)__int_as_float(blk[2]) - (double)__int_as_float(blk[3]);
        }
        loss_out[0] = s + (double)__int_as_float(L.base[3]);
    }
    if (i >= n) return;
    float gi = 0.f;
    for (int r = 0; r < L.world; ++r)
        gi = __fadd_rn(gi, reinterpret_cast<const float*>(L.base + (int64_t)r * L.stride + L.dense_off)[i]);
    g_out[i] = gi;
    const uint8_t f = flags[i];
    if (!(f & 2)) return;
    float var = p[i], mm = m[i], vv = v[i];
    if (f & 1) gi = __fadd_rn(gi, __fmul_rn(hp->reg_lambda, var));
    adam_elem(var, mm, vv, gi, hp->alpha);
    p[i] = var; m[i] = mm; v[i] = vv;
}
void launch_dp_dense_adam(cudaStream_t st, const DpLayout& L, float* p, float* m, float* v, float* g_out,
                          const uint8_t* flags, int n, const Hyper* hp, float* alpha_hist, double* loss_out) {
    dp_dense_adam_kernel<<<(n + 255) / 256, 256, 0, st>>>(L, p, m, v, g_out, flags, n, hp, alpha_hist, loss_out);
    ++g_launch_count;
}

__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

// optimizer step of one row by one group of LPR lanes (all of them active here, identical control flow)
// EXPORT: 0 = apply Adam; 1 = write the run's gradient sum at the run's first sorted index; 2 = at the run's rank among
// the runs of the sorted list (a.head_slot, launch_head_slots): compact AND ascending by key, which is what lets the
// data-parallel finish merge the ranks' lists instead of sorting their concatenation
template <int EXPORT, int LPR>
__device__ __forceinline__ void emb_apply_row(const EmbUpdateArgs& a, int32_t key, int64_t out_idx, const float4& acc,
                                              float4 var, float4 m, float4 v, int last, int sub, float alpha, int step) {
    constexpr int D = LPR * 4;
    if (EXPORT) {
        if (EXPORT == 2 && out_idx >= a.out_cap) return;   // the caller sized the list from the run count: never taken
        *reinterpret_cast<float4*>(a.out_rows + out_idx * D + sub * 4) = acc;
        if (sub == 0) a.out_heads[out_idx] = key;
        return;
    }
    const int64_t off = (int64_t)key * a.es + sub * 4;
    if (a.alpha_hist) {   // LAZY: a row another rank gathered may not be current here yet
        const int upto = step - 1;
        if (last < upto && !(all_zero(m) && all_zero(v))) replay4(var, m, v, last, upto, a.alpha_hist);
    }
    adam4(var, m, v, acc, alpha);
    {   // every lane of the row group has read last_step before lane 0 overwrites it
        const int lane = threadIdx.x & 31;
        const unsigned gmask = (LPR >= 32) ? FULL_MASK : (((1u << LPR) - 1u) << (lane & ~(LPR - 1)));
        __syncwarp(gmask);
    }
    *reinterpret_cast<float4*>(a.emb + off) = var;
    *reinterpret_cast<float4*>(a.m + off) = m;
    *reinterpret_cast<float4*>(a.v + off) = v;
    if (sub == 0 && a.last_step) a.last_step[key] = step;
}

// entries gi, gi+NG, gi+2NG ... of the run, four independent loads at a time, additions in ascending order
template <int LPR>
__device__ __forceinline__ float4 emb_strided_sum(const EmbUpdateArgs& a, int32_t start, int32_t cnt, int gi, int NG, int sub) {
    constexpr int D = LPR * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = gi; e < cnt; e += 4 * NG) {
        int32_t pp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) pp[u] = (e + u * NG < cnt) ? a.spos[(int64_t)start + e + u * NG] : -1;
        float4 g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            g[u] = pp[u] >= 0 ? *reinterpret_cast<const float4*>(a.grad_rows + (int64_t)pp[u] * D + sub * 4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (pp[u] >= 0) add4(acc, g[u]);
    }
    return acc;
}
// sum of the group partials of a team of TL lanes (TL/LPR groups) in group order; every lane of the team ends up with
// the total of its sub-chunk
template <int LPR, int TL>
__device__ __forceinline__ float4 emb_team_combine(const float4& part, int lane) {
    constexpr int GT = TL / LPR;
    const int src0 = (lane & ~(TL - 1)) + (lane % LPR);
    float4 tot;
    tot.x = __shfl_sync(FULL_MASK, part.x, src0); tot.y = __shfl_sync(FULL_MASK, part.y, src0);
    tot.z = __shfl_sync(FULL_MASK, part.z, src0); tot.w = __shfl_sync(FULL_MASK, part.w, src0);
#pragma unroll
    for (int g = 1; g < GT; ++g) {
        float4 o;
        o.x = __shfl_sync(FULL_MASK, part.x, src0 + g * LPR); o.y = __shfl_sync(FULL_MASK, part.y, src0 + g * LPR);
        o.z = __shfl_sync(FULL_MASK, part.z, src0 + g * LPR); o.w = __shfl_sync(FULL_MASK, part.w, src0 + g * LPR);
        add4(tot, o);
    }
    return tot;
}

template <int EXPORT, int LPR>
__global__ void __launch_bounds__(256) emb_update_kernel(EmbUpdateArgs a) {
    pdl_enter();
    constexpr int D = LPR * 4;
    constexpr int NGC = 256 / LPR;       // groups per CTA
    __shared__ float4 red[8][LPR];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = threadIdx.x % LPR;
    const int nS = a.counters[0], nM = a.counters[1], nL = a.counters[2];
    const float alpha = a.hp->alpha;
    const int step = a.hp->step;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int4* __restrict__ runs_long = reinterpret_cast<const int4*>(a.runs_long);

    // ---- tier L: one CTA per run, or per RUN_CHUNK-entry chunk of a longer run (descriptor .w = chunk slot, -1: whole run)
    __shared__ int s_last;
    for (int r = blockIdx.x; r < nL; r += gridDim.x) {
        const int4 dsc = runs_long[a.long_cap - 1 - r];
        const int32_t key = dsc.x, start = dsc.y, cnt = dsc.z, cslot = dsc.w;
        const bool owner = threadIdx.x < LPR;
        float4 var = z4, m = z4, v = z4; int last = 0;
        if (owner && !EXPORT && cslot < 0) {
            const int64_t off = (int64_t)key * a.es + sub * 4;
            var = *reinterpret_cast<const float4*>(a.emb + off);
            m = *reinterpret_cast<const float4*>(a.m + off);
            v = *reinterpret_cast<const float4*>(a.v + off);
            if (a.alpha_hist) last = a.last_step[key];
        }
        const float4 part = emb_strided_sum<LPR>(a, start, cnt, threadIdx.x / LPR, NGC, sub);
        const float4 wtot = emb_team_combine<LPR, 32>(part, lane);
        if (lane < LPR) red[warp][lane] = wtot;
        __syncthreads();
        float4 acc = z4;
        if (owner) {
            acc = red[0][sub];
#pragma unroll
            for (int w = 1; w < 8; ++w) add4(acc, red[w][sub]);
        }
        if (cslot < 0) {
            if (owner)
                emb_apply_row<EXPORT, LPR>(a, key, EXPORT == 2 ? (int64_t)a.head_slot[start] : (int64_t)start, acc, var, m, v, last, sub,
                                           alpha, step);
        } else {
            // chunk of a long run: publish the chunk sum; the CTA that completes the run adds the chunks in chunk order
            const int4 info = reinterpret_cast<const int4*>(a.slotinfo)[cslot];   // {first slot, chunks, run start}
            if (owner) *reinterpret_cast<float4*>(a.part + (int64_t)cslot * D + sub * 4) = acc;
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) s_last = (atomicAdd(a.done + info.x, 1) == info.y - 1) ? 1 : 0;
            __syncthreads();
            if (s_last) {
                __threadfence();
                if (owner) {
                    if (!EXPORT) {
                        const int64_t off = (int64_t)key * a.es + sub * 4;
                        var = *reinterpret_cast<const float4*>(a.emb + off);
                        m = *reinterpret_cast<const float4*>(a.m + off);
                        v = *reinterpret_cast<const float4*>(a.v + off);
                        if (a.alpha_hist) last = a.last_step[key];
                    }
                    const float* pp = a.part + (int64_t)info.x * D + sub * 4;
                    float4 tot = __ldcg(reinterpret_cast<const float4*>(pp));
                    for (int c = 1; c < info.y; ++c) add4(tot, __ldcg(reinterpret_cast<const float4*>(pp + (int64_t)c * D)));
                    emb_apply_row<EXPORT, LPR>(a, key, EXPORT == 2 ? (int64_t)a.head_slot[info.z] : (int64_t)info.z, tot, var, m, v, last,
                                               sub, alpha, step);
                }
                if (threadIdx.x == 0) a.done[info.x] = 0;   // ready for the next launch
            }
        }
        __syncthreads();
    }
    // ---- tier M: one team of TL lanes per run (half a warp for narrow rows: two runs per warp in flight)
    {
        constexpr int TL = LPR <= 8 ? 16 : 32;
        constexpr int TPW = 32 / TL;         // teams per warp
        const int tw = lane / TL, tl = lane % TL;
        const int gwarp = blockIdx.x * 8 + warp, nwarps = gridDim.x * 8;
        for (int r0 = gwarp * TPW; r0 < nM; r0 += nwarps * TPW) {   // warp-uniform trip count
            const int r = r0 + tw;
            const bool valid = r < nM;
            int4 dsc = make_int4(0, 0, 0, 0);
            if (valid) dsc = runs_long[r];
            const int32_t key = dsc.x, start = dsc.y, cnt = dsc.z;
            const bool owner = valid && tl < LPR;
            float4 var = z4, m = z4, v = z4; int last = 0;
            if (owner && !EXPORT) {
                const int64_t off = (int64_t)key * a.es + sub * 4;
                var = *reinterpret_cast<const float4*>(a.emb + off);
                m = *reinterpret_cast<const float4*>(a.m + off);
                v = *reinterpret_cast<const float4*>(a.v + off);
                if (a.alpha_hist) last = a.last_step[key];
            }
            const float4 part = emb_strided_sum<LPR>(a, start, cnt, tl / LPR, TL / LPR, sub);
            __syncwarp();
            const float4 acc = emb_team_combine<LPR, TL>(part, lane);
            if (owner)
                emb_apply_row<EXPORT, LPR>(a, key, EXPORT == 2 ? (int64_t)a.head_slot[start] : (int64_t)start, acc, var, m, v, last, sub,
                                           alpha, step);
            __syncwarp();
        }
    }
    // ---- tier S: one group per run, next descriptor prefetched
    const int ngroups = (int)((gridDim.x * blockDim.x) / LPR);
    const int4* __restrict__ runs = reinterpret_cast<const int4*>(a.runs);
    int h = (int)((blockIdx.x * blockDim.x + threadIdx.x) / LPR);
    if (h >= nS) return;
    int4 d0 = runs[2 * h], d1 = runs[2 * h + 1];
    for (; h < nS; h += ngroups) {
        const int32_t key = d0.x, start = d0.y, n4 = d0.z;
        const int32_t pos[4] = {d0.w, d1.x, d1.y, d1.z};
        {   // clamped: the last iteration re-reads its own descriptor
            const int hn = (h + ngroups < nS) ? h + ngroups : h;
            d0 = runs[2 * hn]; d1 = runs[2 * hn + 1];
        }
        float4 var = z4, m = z4, v = z4; int last = 0;
        if (!EXPORT) {
            const int64_t off = (int64_t)key * a.es + sub * 4;
            var = *reinterpret_cast<const float4*>(a.emb + off);
            m = *reinterpret_cast<const float4*>(a.m + off);
            v = *reinterpret_cast<const float4*>(a.v + off);
            if (a.alpha_hist) last = a.last_step[key];
        }
        float4 g[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            g[u] = (u < n4) ? *reinterpret_cast<const float4*>(a.grad_rows + (int64_t)pos[u] * D + sub * 4) : z4;
        float4 acc = g[0];
#pragma unroll
        for (int u = 1; u < 4; ++u)
            if (u < n4) add4(acc, g[u]);
        emb_apply_row<EXPORT, LPR>(a, key, EXPORT == 2 ? (int64_t)a.head_slot[start] : (int64_t)start, acc, var, m, v, last, sub, alpha, step);
    }
}
template <int LPR>
static void emb_update_launch(cudaStream_t st, const EmbUpdateArgs& a, unsigned grid) {
    if (a.mode == 1) launch_chain(emb_update_kernel<1, LPR>, dim3(grid), dim3(256), 0, st, a);
    else if (a.mode == 2) launch_chain(emb_update_kernel<2, LPR>, dim3(grid), dim3(256), 0, st, a);
    else launch_chain(emb_update_kernel<0, LPR>, dim3(grid), dim3(256), 0, st, a);
}
void launch_emb_update(cudaStream_t st, const EmbUpdateArgs& a) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const int lpr = a.d >> 2;
    // the run counts live on the device: size the grid for one group per sorted index, capped at a few resident waves
    // (groups past the run count exit at once; the grid-stride loops cover the rest)
    static int per_sm = 0;
    if (!per_sm) {
        const char* e = getenv("SCORE_UPD_CTAS_PER_SM");
        per_sm = e ? atoi(e) : 0;
        if (per_sm <= 0) per_sm = 16;
    }
    int64_t want = (a.n * lpr + 255) / 256, cap = (int64_t)sms * per_sm;
    unsigned grid = (unsigned)(want < cap ? want : cap);
    if (grid == 0) grid = 1;
    switch (lpr) {
        case 1: emb_update_launch<1>(st, a, grid); break;
        case 2: emb_update_launch<2>(st, a, grid); break;
        case 4: emb_update_launch<4>(st, a, grid); break;
        case 8: emb_update_launch<8>(st, a, grid); break;
        case 16: emb_update_launch<16>(st, a, grid); break;
        case 32: emb_update_launch<32>(st, a, grid); break;
        default: break;   // score_create rejects other widths
    }
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------ data-parallel update
// Embedding half of the data-parallel finish, ONE kernel over the gathered blocks of all ranks: every rank's id list is
// ascending with one entry per unique row, so no sort and no merged list is needed.  One group of LPR lanes per entry
// (rank r, index i): the lanes split the other ranks among themselves and look the entry's id up in their lists; the
// entry of the LOWEST rank that holds the id owns the row: it adds the gradient rows of the higher ranks in rank order
// (fixed order: every replica computes the same bits) and applies the row's Adam step, the other entries retire.  The
// optimizer state of the row is fetched before the lookups - most entries own their row - so the latencies overlap.
// A lookup is two short binary searches: first in a sampled copy of the list (every DP_SAMPLE-th id; all ranks' samples
// together are a few tens of KB and stay in L1), then inside the one DP_SAMPLE-id block of the full list it points at -
// two or three L2 round trips instead of seventeen; the up-to-DP_J lookups of a lane advance side by side (DP_J = 1, 2 or 4
// lookups per lane, the smallest that covers the world in one round).
constexpr int DP_SAMPLE = 64;

__global__ void dp_sample_kernel(DpLayout L, int ns, int32_t* __restrict__ sample) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)L.world * ns) return;
    const int r = (int)(idx / ns), j = (int)(idx - (int64_t)r * ns);
    const int32_t* blk = L.base + (int64_t)r * L.stride;
    int c = blk[0];
    c = c < 0 ? 0 : (c > L.cap ? (int)L.cap : c);
    const int pos = j * DP_SAMPLE;
    sample[idx] = pos < c ? blk[L.keys_off + pos] : 0x7fffffff;
}

template <int LPR, int DP_J>
__global__ void __launch_bounds__(256) dp_apply_kernel(DpLayout L, int ns, int it1, const int32_t* __restrict__ sample,
                                                        EmbUpdateArgs a, int32_t* __restrict__ err_flag) {
    constexpr int D = LPR * 4;
    const int lane = threadIdx.x & 31;
    const int sub = threadIdx.x % LPR;
    const unsigned gmask = (LPR >= 32) ? FULL_MASK : (((1u << LPR) - 1u) << (lane & ~(LPR - 1)));
    const int glane0 = lane & ~(LPR - 1);
    const int64_t gidx = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    const int r = (int)(gidx / L.cap);
    const int64_t i = gidx - (int64_t)r * L.cap;
    if (r >= L.world) return;
    const int32_t* __restrict__ mine = L.base + (int64_t)r * L.stride;
    int32_t cnt = mine[0];
    if (cnt > L.cap || cnt < 0) { if (i == 0 && sub == 0) atomicExch(err_flag, 2); cnt = cnt < 0 ? 0 : (int32_t)L.cap; }
    if (i >= cnt) return;                                   // uniform over the group
    const int32_t key = mine[L.keys_off + i];
    const int64_t off = (int64_t)key * a.es + sub * 4;
    float4 var = *reinterpret_cast<const float4*>(a.emb + off);
    float4 m = *reinterpret_cast<const float4*>(a.m + off);
    float4 v = *reinterpret_cast<const float4*>(a.v + off);
    const int last = a.alpha_hist ? a.last_step[key] : 0;
    float4 acc = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(mine + L.keys_off + L.cap) + i * D + sub * 4);
    for (int r0 = 0; r0 < L.world; r0 += LPR * DP_J) {
        // lane `sub` looks the id up in the lists of ranks r0 + sub, r0 + LPR + sub, ... (DP_J lookups side by side)
        const int32_t* kp[DP_J];
        const int32_t* sp[DP_J];
        int lo[DP_J], hi[DP_J], c[DP_J], found[DP_J];
#pragma unroll
        for (int j = 0; j < DP_J; ++j) {
            const int rr = r0 + j * LPR + sub;
            const bool valid = rr < L.world && rr != r;
            const int32_t* blk = L.base + (int64_t)(valid ? rr : r) * L.stride;
            kp[j] = blk + L.keys_off;
            sp[j] = sample + (int64_t)(valid ? rr : r) * ns;
            int cc = valid ? blk[0] : 0;
            c[j] = cc < 0 ? 0 : (cc > L.cap ? (int)L.cap : cc);
            lo[j] = 0; hi[j] = c[j] > 0 ? ns : 0;           // phase 1: first sample > key (samples past the count are INT_MAX)
            found[j] = -1;
        }
        for (int it = 0; it < it1; ++it) {
#pragma unroll
            for (int j = 0; j < DP_J; ++j)
                if (lo[j] < hi[j]) {
                    const int mid = (lo[j] + hi[j]) >> 1;
                    if (__ldg(sp[j] + mid) <= key) lo[j] = mid + 1; else hi[j] = mid;
                }
        }
#pragma unroll
        for (int j = 0; j < DP_J; ++j) {                    // phase 2: lower bound inside block lo-1 of the full list
            const int blk0 = (lo[j] - 1) * DP_SAMPLE;
            const bool any = c[j] > 0 && lo[j] > 0;
            const int end = blk0 + DP_SAMPLE < c[j] ? blk0 + DP_SAMPLE : c[j];
            lo[j] = any ? blk0 : 0; hi[j] = any ? end : 0;
        }
#pragma unroll
        for (int it = 0; it < 7; ++it) {                    // 2^6 = DP_SAMPLE ids: at most 7 halvings
#pragma unroll
            for (int j = 0; j < DP_J; ++j)
                if (lo[j] < hi[j]) {
                    const int mid = (lo[j] + hi[j]) >> 1;
                    if (kp[j][mid] < key) lo[j] = mid + 1; else hi[j] = mid;
                }
        }
#pragma unroll
        for (int j = 0; j < DP_J; ++j)
            if (lo[j] < c[j] && c[j] > 0 && kp[j][lo[j]] == key) found[j] = lo[j];
#pragma unroll
        for (int j = 0; j < DP_J; ++j) {
#pragma unroll
            for (int q = 0; q < LPR; ++q) {
                const int rq = r0 + j * LPR + q;
                if (rq >= L.world) continue;                // uniform
                const int fq = __shfl_sync(gmask, found[j], glane0 + q);
                if (fq < 0) continue;
                if (rq < r) return;                         // a lower rank owns the row (uniform over the group)
                const float* rows_q = reinterpret_cast<const float*>(L.base + (int64_t)rq * L.stride + L.keys_off + L.cap);
                const float4 g = *reinterpret_cast<const float4*>(rows_q + (int64_t)fq * D + sub * 4);
                acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
            }
        }
    }
    emb_apply_row<0, LPR>(a, key, 0, acc, var, m, v, last, sub, a.hp->alpha, a.hp->step);
}
int64_t dp_sample_count(int world, int64_t cap) { return (int64_t)world * (cap / DP_SAMPLE); }
void launch_dp_apply(cudaStream_t st, const DpLayout& L, const EmbUpdateArgs& a, int32_t* sample, int32_t* err_flag) {
    const int lpr = a.d >> 2;
    const int ns = (int)(L.cap / DP_SAMPLE);                // cap is a multiple of 1024
    dp_sample_kernel<<<(unsigned)(((int64_t)L.world * ns + 255) / 256), 256, 0, st>>>(L, ns, sample);
    const int64_t threads = (int64_t)L.world * L.cap * lpr;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    int it1 = 1;
    while ((1 << it1) <= ns) ++it1;
#define DP_APPLY(LPR_)                                                                                              \
    if (L.world <= (LPR_) + 1) dp_apply_kernel<LPR_, 1><<<grid, 256, 0, st>>>(L, ns, it1, sample, a, err_flag);       \
    else if (L.world <= 2 * (LPR_)) dp_apply_kernel<LPR_, 2><<<grid, 256, 0, st>>>(L, ns, it1, sample, a, err_flag);  \
    else dp_apply_kernel<LPR_, 4><<<grid, 256, 0, st>>>(L, ns, it1, sample, a, err_flag)
    switch (lpr) {
        case 1: DP_APPLY(1); break;
        case 2: DP_APPLY(2); break;
        case 4: DP_APPLY(4); break;
        case 8: DP_APPLY(8); break;
        case 16: DP_APPLY(16); break;
        case 32: DP_APPLY(32); break;
        default: break;
    }
#undef DP_APPLY
    g_launch_count += 2;
}

// DENSE mode: the zero-gradient Adam step of every row the batch did not touch (row 0 included:
// its gradient is always zero, so its slots stay zero and it never moves).
__global__ void emb_dense_sweep_kernel(float* __restrict__ emb, float* __restrict__ m, float* __restrict__ v,
                                       const int32_t* __restrict__ last_step, int64_t V, int d, int es, const Hyper* hp) {
    const int lpr = d >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= V * lpr) return;
    const int64_t row = idx / lpr;
    const int64_t off = row * es + (idx - row * lpr) * 4;
    if (last_step[row] == hp->step) return;
    float4 mm = *reinterpret_cast<float4*>(m + off);
    float4 vv = *reinterpret_cast<float4*>(v + off);
    if (all_zero(mm) && all_zero(vv)) return;   // never touched: the update is exactly a no-op
    float4 var = *reinterpret_cast<float4*>(emb + off);
    adam4(var, mm, vv, make_float4(0.f, 0.f, 0.f, 0.f), hp->alpha);
    *reinterpret_cast<float4*>(emb + off) = var;
    *reinterpret_cast<float4*>(m + off) = mm;
    *reinterpret_cast<float4*>(v + off) = vv;
}
void launch_emb_dense_sweep(cudaStream_t st, float* emb, float* m, float* v, int32_t* last_step, int64_t V, int d, int es,
                            const Hyper* hp) {
    int64_t n = V * (d >> 2);
    emb_dense_sweep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(emb, m, v, last_step, V, d, es, hp);
    ++g_launch_count;
}

// LAZY catch-up of the rows about to be gathered, in two kernels:
//   emb_claim_kernel   one thread per gathered position; the thread that wins the atomic claim of a stale row appends
//                      (row, last step) to a compact list (plain pre-check first: duplicates and current rows skip the atomic);
//   emb_replay_kernel  one thread per ELEMENT of a claimed row replays the skipped zero-gradient steps.
// The replay is a serial chain of IEEE div/sqrt per element, i.e. instruction-bound; a group-per-position layout leaves
// the lanes of unclaimed positions idle and makes a warp wait for the longest of its eight replays.  The compact list
// keeps every lane busy and only two rows share a warp.  The replay result does not depend on which duplicate wins.
__global__ void emb_claim_kernel(const int32_t* __restrict__ keys, int64_t n, int64_t V, int32_t* __restrict__ last_step,
                                 const Hyper* hp, int32_t* __restrict__ list, int32_t* __restrict__ counter) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int upto = hp->step - 1;
    const int32_t key = (i < n) ? keys[i] : 0;
    bool win = false;
    int old = upto;
    if (key > 0 && (int64_t)key < V) {
        const int seen = last_step[key];
        if ((uint32_t)seen < (uint32_t)upto) {   // -1: never touched by an optimizer step, nothing to replay (build_keys_kernel)
            old = atomicExch(&last_step[key], upto);
            win = old < upto;
        }
    }
    const unsigned mask = __ballot_sync(FULL_MASK, win);
    if (mask == 0) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(FULL_MASK, base, 0);
    if (win) {
        const int idx = base + __popc(mask & ((1u << lane) - 1u));
        list[2 * idx] = key;
        list[2 * idx + 1] = old;
    }
}
__global__ void __launch_bounds__(256) emb_replay_kernel(const int32_t* __restrict__ list, const int32_t* __restrict__ counter,
                                                         float* __restrict__ emb, float* __restrict__ m, float* __restrict__ v,
                                                         int d, int es, const float* __restrict__ alpha_hist, const Hyper* hp, int pingpong) {
    pdl_enter();
    const int upto = hp->step - 1;
    const int64_t total = (int64_t)(counter[pingpong ? (hp->seq & 1) : 0]) * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / d;
        const int c = (int)(e - r * d);
        const int32_t key = list[2 * r], old = list[2 * r + 1];
        const int64_t off = (int64_t)key * es + c;
        float mm = m[off], vv = v[off];
        if (mm == 0.f && vv == 0.f) continue;   // never touched: every replayed step is exactly a no-op
        float var = emb[off];
        for (int s = old + 1; s <= upto; ++s) adam_elem(var, mm, vv, 0.f, alpha_hist[s]);
        emb[off] = var; m[off] = mm; v[off] = vv;
    }
}
static int num_sms_cached() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}
void launch_emb_replay(cudaStream_t st, const int32_t* claim_list, const int32_t* claim_counter, int64_t max_rows, float* emb,
                       float* m, float* v, int d, int es, const float* alpha_hist, const Hyper* hp, int pingpong) {
    int64_t want = (max_rows * d + 255) / 256, cap = (int64_t)num_sms_cached() * 8;
    if (want < 1) want = 1;
    launch_chain(emb_replay_kernel, dim3((unsigned)(want < cap ? want : cap)), dim3(256), 0, st, claim_list, claim_counter, emb, m, v, d, es, alpha_hist, hp, pingpong);
    ++g_launch_count;
}
void launch_emb_catchup_rows(cudaStream_t st, const int32_t* keys, int64_t n, int64_t V, float* emb, float* m, float* v,
                             int32_t* last_step, int d, int es, const float* alpha_hist, const Hyper* hp, int32_t* claim_list,
                             int32_t* claim_counter) {
    cudaMemsetAsync(claim_counter, 0, sizeof(int32_t), st);
    emb_claim_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys, n, V, last_step, hp, claim_list, claim_counter);
    ++g_launch_count;
    launch_emb_replay(st, claim_list, claim_counter, n, emb, m, v, d, es, alpha_hist, hp);
    cudaMemsetAsync(claim_counter, 0, 2 * sizeof(int32_t), st);   // leave the ping-pong pair of the fused path clean
}
__global__ void emb_catchup_all_kernel(float* __restrict__ emb, float* __restrict__ m, float* __restrict__ v,
                                       int32_t* __restrict__ last_step, int64_t V, int d, int es,
                                       const float* __restrict__ alpha_hist, int upto) {
    const int lpr = d >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = idx / lpr;
    const bool ok = idx < V * lpr;
    int old = ok ? last_step[row] : upto;
    // every lane of a row must read last_step before lane 0 overwrites it
    __syncwarp();
    if (!ok || old >= upto) return;
    const int64_t off = row * es + (idx - row * lpr) * 4;
    float4 mm = *reinterpret_cast<float4*>(m + off);
    float4 vv = *reinterpret_cast<float4*>(v + off);
    if (!(all_zero(mm) && all_zero(vv))) {
        float4 var = *reinterpret_cast<float4*>(emb + off);
        replay4(var, mm, vv, old, upto, alpha_hist);
        *reinterpret_cast<float4*>(emb + off) = var;
        *reinterpret_cast<float4*>(m + off) = mm;
        *reinterpret_cast<float4*>(v + off) = vv;
    }
}
__global__ void set_last_step_kernel(int32_t* last_step, int64_t V, int upto) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V && (uint32_t)last_step[i] < (uint32_t)upto) last_step[i] = upto;   // the never-touched mark (-1) stays
}
void launch_emb_catchup_all(cudaStream_t st, float* emb, float* m, float* v, int32_t* last_step, int64_t V, int d, int es,
                            const float* alpha_hist, int upto) {
    int64_t n = V * (d >> 2);
    emb_catchup_all_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(emb, m, v, last_step, V, d, es, alpha_hist, upto);
    set_last_step_kernel<<<(unsigned)((V + 255) / 256), 256, 0, st>>>(last_step, V, upto);
    g_launch_count += 2;
}

// ------------------------------------------------------------------------------------------ plain row gather
__global__ void gather_rows_kernel(const float* __restrict__ table, int es, const int32_t* __restrict__ idx, int64_t n, int d,
                                   int64_t V, float* __restrict__ out, int32_t* __restrict__ err_flag) {
    const int lpr = d >> 2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = t / lpr;
    const int sub = (int)(t - i * lpr);
    if (i >= n) return;
    int32_t id = idx[i];
    if (id < 0 || id >= V) { atomicExch(err_flag, 1); id = 0; }
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (id != 0) v = *reinterpret_cast<const float4*>(table + (int64_t)id * es + sub * 4);
    *reinterpret_cast<float4*>(out + i * d + sub * 4) = v;
}
void launch_gather_rows(cudaStream_t st, const float* table, int es, const int32_t* idx, int64_t n, int d, int64_t V, float* out,
                        int32_t* err_flag) {
    const int64_t threads = n * (d >> 2);
    gather_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(table, es, idx, n, d, V, out, err_flag);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------ initialisers
__global__ void init_trunc_normal_kernel(float* __restrict__ p, int64_t n, uint32_t lo, uint32_t hi, uint32_t stream_id,
                                         int row_len, int row_stride) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float z = 0.f;
    for (uint32_t attempt = 0; attempt < 64; ++attempt) {   // tf.truncated_normal: re-draw beyond 2 sigma
        float u1 = philox_uniform(lo, hi, stream_id, 2 * attempt, (uint64_t)i);
        float u2 = philox_uniform(lo, hi, stream_id, 2 * attempt + 1, (uint64_t)i);
        u1 = fmaxf(u1, 5.9604645e-8f);
        z = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
        if (fabsf(z) <= 2.0f) break;
        z = 0.f;
    }
    // the random stream is keyed by the LOGICAL element index, so the values do not depend on the row layout
    const int64_t at = row_len > 0 ? (i / row_len) * row_stride + i % row_len : i;
    p[at] = z;
}
void launch_init_trunc_normal(cudaStream_t st, float* p, int64_t n, uint64_t seed, uint32_t stream_id, int row_len,
                              int row_stride) {
    init_trunc_normal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                                          stream_id, row_len, row_stride);
    ++g_launch_count;
}
__global__ void init_uniform_kernel(float* __restrict__ p, int64_t n, float limit, uint32_t lo, uint32_t hi,
                                    uint32_t stream_id) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float u = philox_uniform(lo, hi, stream_id, 0, (uint64_t)i);
    p[i] = (2.0f * u - 1.0f) * limit;
}
void launch_init_uniform(cudaStream_t st, float* p, int64_t n, float limit, uint64_t seed, uint32_t stream_id) {
    init_uniform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, limit, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                                     stream_id);
    ++g_launch_count;
}
__global__ void fill_kernel(float* p, int64_t n, float v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
void launch_fill(cudaStream_t st, float* p, int64_t n, float v) {
    if (n <= 0) return;
    fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, v);
    ++g_launch_count;
}
__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
void launch_fill_i32(cudaStream_t st, int32_t* p, int64_t n, int32_t v) {
    if (n <= 0) return;
    fill_i32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, v);
    ++g_launch_count;
}

}  // namespace score
