// Embedding front end of the SCoRe path: fused gather + cross-neighbor co-attention + pooling.
//
// Replaces, without materialising anything the reference materialises:
//   * emb_mtx * emb_mtx_mask                      (score.py:45-47)  -> id 0 is zero-filled in the gather
//   * six tf.nn.embedding_lookup + reshape         (score.py:51-66)  -> cp.async 16-byte row chunks into smem
//   * co_attention()                               (score.py:147-167)-> rank-1 form, see below
//
// co_attention tiles BOTH seq1 and seq2 along axis 3 (score.py:152-153), so
//   rel[b,t,i,j] = relu(Wt.target + W1.seq1[i] + W2.seq2[i] + bias) =: r_i        (independent of j)
//   seq1_weights[i] = softmax_i(r_i)        seq2_weights[j] = 1/K
//   atten_info = [K*r_i (K values) || (sum_i r_i) repeated K times]
// One warp owns one (b, t) time slice: it stages the slice's K*(2*if + 2*uf) table rows in shared
// memory with cp.async (HBM-bound phase, many slices in flight per SM), then does the dot
// products with warp-shuffle reductions.  Slices with t >= length[b] are skipped: they reach
// neither the GRU output nor the attention (score.py:179-181, 205-208).
#include "kernels.h"

namespace score {

// ------------------------------------------------------------------------------------------
// keys[p] = sanitized id of flat position p: 0 for masked slices (t >= length[b]) and ids that
// are out of range (flagged).  Positions: [user_1hop | user_2hop | item_1hop | item_2hop | tu | ti].
// Four positions per thread (one 16-byte id load, four independent dependent chains).  In LAZY optimizer mode the
// same pass claims the stale rows among the keys (atomic exchange on last_step, plain pre-check first) and appends
// (row, last step) to the compact list emb_replay_kernel works through (scatter.cu).
__device__ __forceinline__ bool position_live(const Dims& dm, int64_t p, const int32_t* __restrict__ length) {
    if (p >= dm.off_tu) return true;
    int64_t local; int f;
    if (p < dm.off_u2) { local = p - dm.off_u1; f = dm.fi; }
    else if (p < dm.off_i1) { local = p - dm.off_u2; f = dm.fu; }
    else if (p < dm.off_i2) { local = p - dm.off_i1; f = dm.fu; }
    else { local = p - dm.off_i2; f = dm.fi; }
    const int64_t slice = local / ((int64_t)dm.K * f);
    const int b = (int)(slice / dm.T), t = (int)(slice % dm.T);
    return t < length[b];
}

__global__ void build_keys_kernel(Dims dm, const int32_t* __restrict__ ids, const int32_t* __restrict__ length,
                                  int32_t* __restrict__ keys, int32_t* __restrict__ err_flag, ClaimArgs ca) {
    const int64_t p0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int lane = threadIdx.x & 31;
    int32_t id[4] = {0, 0, 0, 0};
    if (p0 + 3 < dm.N) {
        const int4 v = *reinterpret_cast<const int4*>(ids + p0);
        id[0] = v.x; id[1] = v.y; id[2] = v.z; id[3] = v.w;
    } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) if (p0 + u < dm.N) id[u] = ids[p0 + u];
    }
    bool bad = false;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (p0 + u >= dm.N) { id[u] = 0; continue; }
        if (id[u] < 0 || (int64_t)id[u] >= dm.V) { bad = true; id[u] = 0; }
        if (id[u] != 0 && !position_live(dm, p0 + u, length)) id[u] = 0;
    }
    if (bad) atomicExch(err_flag, 1);
    if (p0 + 3 < dm.N) *reinterpret_cast<int4*>(keys + p0) = make_int4(id[0], id[1], id[2], id[3]);
    else {
#pragma unroll
        for (int u = 0; u < 4; ++u) if (p0 + u < dm.N) keys[p0 + u] = id[u];
    }
    if (!ca.last_step) return;   // uniform
    const int upto = ca.hp->step - 1;
    int seen[4], old[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) seen[u] = id[u] != 0 ? ca.last_step[id[u]] : upto;   // four independent loads
    int nwin = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        old[u] = upto;
        if (seen[u] < upto) old[u] = atomicExch(&ca.last_step[id[u]], upto);
        nwin += old[u] < upto ? 1 : 0;
    }
    // warp-aggregated append
    int incl = nwin;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += y;
    }
    const int total = __shfl_sync(FULL_MASK, incl, 31);
    if (total == 0) return;
    int base = 0;
    if (lane == 31) base = atomicAdd(ca.counter, total);
    base = __shfl_sync(FULL_MASK, base, 31);
    int idx = base + incl - nwin;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (old[u] < upto) { ca.list[2 * idx] = id[u]; ca.list[2 * idx + 1] = old[u]; ++idx; }
}

void launch_build_keys(cudaStream_t st, const Dims& dm, const int32_t* ids, const int32_t* length,
                       int32_t* keys, int32_t* err_flag, const ClaimArgs* claim) {
    ClaimArgs ca{};
    if (claim) { ca = *claim; cudaMemsetAsync(ca.counter, 0, sizeof(int32_t), st); }
    const int threads = 256;
    const int64_t blocks = ((dm.N + 3) / 4 + threads - 1) / threads;
    build_keys_kernel<<<(unsigned)blocks, threads, 0, st>>>(dm, ids, length, keys, err_flag, ca);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------
// stage `nrows` table rows (ids at key_ptr[0..nrows), global or shared) into smem dst[nrows][d]; whole warp
// cooperates, consecutive lanes fetch consecutive 16-byte chunks (a d=16 row is 4 lanes, 64 B contiguous).
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ emb, const int32_t* key_ptr,
                                           int nrows, int d, int lane) {
    const int cpr = d >> 2;   // 16-byte chunks per row
    const int total = nrows * cpr;
    for (int q = lane; q < total; q += 32) {
        int r = q / cpr, c = q - r * cpr;
        int32_t id = key_ptr[r];
        const float* src = emb + (int64_t)id * d + c * 4;
        cp_async16(dst + r * d + c * 4, id != 0 ? src : emb, id != 0 ? 16 : 0);
    }
}

// ------------------------------------------------------------------------------------------
// target rows: one warp per sample b
__global__ void target_fwd_kernel(Dims dm, TargetArgs a) {
    extern __shared__ float sm[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* buf = sm + warp * dm.Ds;   // [tu (Du) | ti (Di)]
    int b = blockIdx.x * warps + warp;
    if (b >= dm.B) return;
    stage_rows(buf, a.emb, a.keys + dm.off_tu + (int64_t)b * dm.fu, dm.fu, dm.d, lane);
    stage_rows(buf + dm.Du, a.emb, a.keys + dm.off_ti + (int64_t)b * dm.fi, dm.fi, dm.d, lane);
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    const float* tu = buf; const float* ti = buf + dm.Du;
    float pu = 0.f, pi = 0.f;
    for (int c = lane; c < dm.Du; c += 32) {
        float v = tu[c];
        a.q0[(int64_t)b * dm.Ds + c] = v;
        if (a.fc_in) a.fc_in[(int64_t)b * dm.Dfc + a.fc_off + dm.Di + c] = v;
        if (a.w_user) pu += a.w_user[c] * v;
    }
    for (int c = lane; c < dm.Di; c += 32) {
        float v = ti[c];
        a.q0[(int64_t)b * dm.Ds + dm.Du + c] = v;
        if (a.fc_in) a.fc_in[(int64_t)b * dm.Dfc + a.fc_off + c] = v;
        if (a.w_item) pi += a.w_item[c] * v;
    }
    pu = warp_sum(pu); pi = warp_sum(pi);
    if (lane == 0 && a.w_item) { a.c_item[b] = pi + a.b_item[0]; a.c_user[b] = pu + a.b_user[0]; }
}

void launch_target_fwd(cudaStream_t st, const Dims& dm, const TargetArgs& a) {
    const int warps = 4;
    size_t smem = (size_t)warps * dm.Ds * sizeof(float);
    target_fwd_kernel<<<(dm.B + warps - 1) / warps, warps * 32, smem, st>>>(dm, a);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------
// Co-attention kernels.  One warp owns one (b, t) slice.
//
// Shared memory per warp: the slice's table rows (16-byte chunks, XOR-swizzled by row so that a lane-per-row
// read of the same chunk index is bank-conflict free), its ids, one dot product per row, the K softmax
// weights / relatedness gradients of both co-attentions, and (backward) the slice's incoming gradients and
// the warp's accumulators for the co-attention kernel gradient.
//
// Row order in shared memory:  seg0 = user_1hop (K*fi rows) | seg1 = item_2hop (K*fi) | seg2 = user_2hop (K*fu)
// | seg3 = item_1hop (K*fu).  Co-attention #1 = (seg0, seg1, target_item), #2 = (seg2, seg3, target_user)
// (score.py:196-197).  The instruction budget matters as much as the bytes here (ncu: the first version
// of this kernel was issue-bound at 2 750 warp instructions per slice), hence lane-per-row float4 dot
// products and lane-per-chunk pooling instead of lane-per-element loops with a shuffle tree per neighbor.
struct CoattSmem {
    int w_off;        // [3*Di + 3*Du] co-attention kernels (CTA-shared)
    int warp_off;     // per-warp region start (floats)
    int warp_stride;  // per-warp floats
    int rows;         // nrows * d
    int nrows;        // K * (2*fi + 2*fu)
    // per-warp layout (floats): rows | kbuf[nrows] | dots[nrows] | wts[4*K pad 4*32] | (bwd) dbuf[2*Ds + 4K] | (bwd) acc[2*Di+2*Du]
    int kbuf_off, dots_off, wts_off, dbuf_off, acc_off;
};
__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }

struct SliceGeom {
    int nfi, nfu, nrows, cpr, swz;
};
__device__ __forceinline__ SliceGeom slice_geom(const Dims& dm) {
    SliceGeom g;
    g.nfi = dm.K * dm.fi; g.nfu = dm.K * dm.fu; g.nrows = 2 * g.nfi + 2 * g.nfu;
    g.cpr = dm.d >> 2; g.swz = min(g.cpr - 1, 7);
    return g;
}
// physical float offset of 16-byte chunk c4 of row r
__device__ __forceinline__ int chunk_off(const SliceGeom& g, int r, int c4) { return (r * g.cpr + (c4 ^ (r & g.swz))) << 2; }

// Gather one (b,t) slice: first ALL of its ids in one batch of independent loads (one global latency), then
// all of its table rows with cp.async (a second latency) - never an id load in front of each row.
__device__ __forceinline__ void stage_slice(float* rows, int32_t* kbuf, const Dims& dm, const SliceGeom& g,
                                            const float* __restrict__ emb, const int32_t* __restrict__ keys,
                                            int64_t slice, int lane) {
    const int32_t* g0 = keys + dm.off_u1 + slice * g.nfi;
    const int32_t* g1 = keys + dm.off_i2 + slice * g.nfi;
    const int32_t* g2 = keys + dm.off_u2 + slice * g.nfu;
    const int32_t* g3 = keys + dm.off_i1 + slice * g.nfu;
    for (int r = lane; r < g.nfi; r += 32) { kbuf[r] = __ldg(g0 + r); kbuf[g.nfi + r] = __ldg(g1 + r); }
    for (int r = lane; r < g.nfu; r += 32) { kbuf[2 * g.nfi + r] = __ldg(g2 + r); kbuf[2 * g.nfi + g.nfu + r] = __ldg(g3 + r); }
    __syncwarp();
    const int total = g.nrows * g.cpr;
    const int shift = 31 - __clz(g.cpr);   // cpr is a power of two
    for (int q = lane; q < total; q += 32) {
        const int r = q >> shift, c4 = q & (g.cpr - 1);
        const int32_t id = kbuf[r];
        const float* src = emb + (int64_t)id * dm.d + c4 * 4;
        cp_async16(rows + chunk_off(g, r, c4), id != 0 ? src : emb, id != 0 ? 16 : 0);
    }
    cp_async_commit();
}

// segment of a row and its position inside it
__device__ __forceinline__ void row_decode(const SliceGeom& g, const Dims& dm, int r, int& seg, int& i, int& field) {
    int rl, f;
    if (r < g.nfi) { seg = 0; rl = r; f = dm.fi; }
    else if (r < 2 * g.nfi) { seg = 1; rl = r - g.nfi; f = dm.fi; }
    else if (r < 2 * g.nfi + g.nfu) { seg = 2; rl = r - 2 * g.nfi; f = dm.fu; }
    else { seg = 3; rl = r - 2 * g.nfi - g.nfu; f = dm.fu; }
    i = rl / f; field = rl - i * f;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// softmax over the first K lanes of `r`; returns the weight of this lane (0 for lanes >= K)
__device__ __forceinline__ float softmax_k(float r, int lane, int K) {
    float x = (lane < K) ? r : -INFINITY;
    float mx = warp_max(x);
    float e = (lane < K) ? expf(x - mx) : 0.f;
    return e / warp_sum(e);
}

__global__ void coatt_fwd_kernel(Dims dm, CoattArgs a, CoattSmem sp) {
    extern __shared__ __align__(16) float sm[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Wsm = sm + sp.w_off;
    for (int i = threadIdx.x; i < 3 * dm.Di; i += blockDim.x) Wsm[i] = a.w_item[i];
    for (int i = threadIdx.x; i < 3 * dm.Du; i += blockDim.x) Wsm[3 * dm.Di + i] = a.w_user[i];
    __syncthreads();
    float* rows = sm + sp.warp_off + warp * sp.warp_stride;
    int32_t* kbuf = reinterpret_cast<int32_t*>(rows + sp.kbuf_off);
    float* dots = rows + sp.dots_off;
    float* wts = rows + sp.wts_off;
    const SliceGeom g = slice_geom(dm);
    const int K = dm.K, Di = dm.Di, Du = dm.Du, Ds = dm.Ds, d = dm.d;
    const float invK = 1.0f / (float)K;
    const int64_t M = (int64_t)dm.B * dm.T;
    for (int64_t slice = (int64_t)blockIdx.x * warps + warp; slice < M; slice += (int64_t)gridDim.x * warps) {
        const int b = (int)(slice / dm.T), t = (int)(slice - (int64_t)b * dm.T);
        float* xu_g = a.xhg_u + slice * dm.ldx; float* xu_c = a.xhc_u + slice * dm.ldx;
        float* xi_g = a.xhg_i + slice * dm.ldx; float* xi_c = a.xhc_i + slice * dm.ldx;
        float* info = a.key + slice * a.ldkey + a.key_off;
        if (t >= a.length[b]) {   // dead slice: nothing downstream reads it, keep buffers finite
            for (int c = lane; c < Ds; c += 32) { xu_g[c] = 0.f; xu_c[c] = 0.f; xi_g[c] = 0.f; xi_c[c] = 0.f; }
            for (int c = lane; c < 4 * K; c += 32) info[c] = 0.f;
            continue;
        }
        stage_slice(rows, kbuf, dm, g, a.emb, a.keys, slice, lane);
        cp_async_wait<0>();
        __syncwarp();
        // (1) one dot product per row with its slice of the co-attention kernel (W1 for seq1 rows, W2 for seq2 rows)
        for (int r = lane; r < g.nrows; r += 32) {
            int seg, i, field;
            row_decode(g, dm, r, seg, i, field);
            const int woff = (seg == 0 ? Di : seg == 1 ? 2 * Di : seg == 2 ? 3 * Di + Du : 3 * Di + 2 * Du) + field * d;
            const float4* wv = reinterpret_cast<const float4*>(Wsm + woff);
            float acc = 0.f;
            for (int c4 = 0; c4 < g.cpr; ++c4)
                acc += dot4(*reinterpret_cast<const float4*>(rows + chunk_off(g, r, c4)), wv[c4]);
            dots[r] = acc;
        }
        __syncwarp();
        // (2) relatedness r_i = relu(target part + seq1[i] part + seq2[i] part), softmax over the K neighbors
        float z1 = a.c_item[b], z2 = a.c_user[b];
        if (lane < K) {
            for (int f = 0; f < dm.fi; ++f) z1 += dots[lane * dm.fi + f] + dots[g.nfi + lane * dm.fi + f];
            for (int f = 0; f < dm.fu; ++f) z2 += dots[2 * g.nfi + lane * dm.fu + f] + dots[2 * g.nfi + g.nfu + lane * dm.fu + f];
        }
        const float r1 = (lane < K) ? fmaxf(z1, 0.f) : 0.f, r2 = (lane < K) ? fmaxf(z2, 0.f) : 0.f;
        const float w1 = softmax_k(r1, lane, K), w2 = softmax_k(r2, lane, K);
        const float s1 = warp_sum(r1), s2 = warp_sum(r2);
        if (lane < K) {
            wts[lane] = w1; wts[K + lane] = w2;
            info[lane] = (float)K * r1; info[K + lane] = s1;            // atten_info of co-attention #1 (score.py:165-166)
            info[2 * K + lane] = (float)K * r2; info[3 * K + lane] = s2;   // ... of co-attention #2
            float* sr = a.save_r + slice * 2 * K; float* sw = a.save_w + slice * 2 * K;
            sr[lane] = r1; sr[K + lane] = r2; sw[lane] = w1; sw[K + lane] = w2;
        }
        __syncwarp();
        // (3) pooling, one 16-byte output chunk per lane:
        //   user_side = [sum_i w1_i user_1hop[i] | sum_i w2_i user_2hop[i]],  item_side = [mean item_1hop | mean item_2hop]
        const int nchunk = (2 * Ds) >> 2;
        for (int e4 = lane; e4 < nchunk; e4 += 32) {
            const int e = e4 << 2;
            int segbase, f, c, wsel; float* dst0; float* dst1;
            if (e < Di) { segbase = 0; f = dm.fi; c = e; wsel = 0; dst0 = xu_g + c; dst1 = xu_c + c; }
            else if (e < Ds) { segbase = 2 * g.nfi; f = dm.fu; c = e - Di; wsel = 1; dst0 = xu_g + Di + c; dst1 = xu_c + Di + c; }
            else if (e < Ds + Du) { segbase = 2 * g.nfi + g.nfu; f = dm.fu; c = e - Ds; wsel = 2; dst0 = xi_g + c; dst1 = xi_c + c; }
            else { segbase = g.nfi; f = dm.fi; c = e - Ds - Du; wsel = 2; dst0 = xi_g + Du + c; dst1 = xi_c + Du + c; }
            const int field = c / d, c4 = (c - field * d) >> 2;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = 0; i < K; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(rows + chunk_off(g, segbase + i * f + field, c4));
                const float w = wsel == 2 ? invK : wts[wsel * K + i];
                acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
            }
            dst0[0] = acc.x; dst0[1] = acc.y; dst0[2] = acc.z; dst0[3] = acc.w;
            dst1[0] = acc.x; dst1[1] = acc.y; dst1[2] = acc.z; dst1[3] = acc.w;
        }
        __syncwarp();
    }
}

static CoattSmem coatt_plan(const Dims& dm, int warps, bool bwd, size_t* bytes) {
    CoattSmem sp;
    sp.w_off = 0;
    sp.warp_off = round4(3 * dm.Di + 3 * dm.Du);
    sp.nrows = dm.K * (2 * dm.fi + 2 * dm.fu);
    sp.rows = sp.nrows * dm.d;
    sp.kbuf_off = sp.rows;
    sp.dots_off = sp.kbuf_off + round4(sp.nrows);
    sp.wts_off = sp.dots_off + round4(sp.nrows);
    sp.dbuf_off = sp.wts_off + 4 * 32;
    sp.acc_off = sp.dbuf_off + (bwd ? round4(2 * dm.Ds + 4 * dm.K) : 0);
    sp.warp_stride = sp.acc_off + (bwd ? round4(2 * dm.Di + 2 * dm.Du) : 0);
    *bytes = (size_t)(sp.warp_off + warps * sp.warp_stride) * sizeof(float);
    return sp;
}

static int g_num_sms = 0;
static int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

void launch_coatt_fwd(cudaStream_t st, const Dims& dm, const CoattArgs& a) {
    const int warps = 4;
    size_t smem;
    CoattSmem sp = coatt_plan(dm, warps, false, &smem);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaFuncSetAttribute(coatt_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = smem;
    }
    int64_t M = (int64_t)dm.B * dm.T;
    int64_t want = (M + warps - 1) / warps;
    int64_t cap = (int64_t)num_sms() * 32;
    int grid = (int)(want < cap ? want : cap);
    coatt_fwd_kernel<<<grid, warps * 32, smem, st>>>(dm, a, sp);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------
// backward of both co-attentions of one slice (same staging as forward; the rows are re-gathered, mostly from L2,
// instead of being stored by the forward pass):
//   dw_i = dout1 . seq1[i]                    dr_i = K*dinfo[i] + sum_j dinfo[K+j] + w_i (dw_i - sum_k w_k dw_k)
//   dz_i = dr_i [r_i > 0]                     dseq1[i] = w_i dout1 + dz_i W1      dseq2[i] = dout2 / K + dz_i W2
//   dW1 += sum_i dz_i seq1[i]                 dW2 += sum_i dz_i seq2[i]           sdz = sum_i dz_i
__global__ void coatt_bwd_kernel(Dims dm, CoattBwdArgs a, CoattSmem sp) {
    extern __shared__ __align__(16) float sm[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Wsm = sm + sp.w_off;
    for (int i = threadIdx.x; i < 3 * dm.Di; i += blockDim.x) Wsm[i] = a.w_item[i];
    for (int i = threadIdx.x; i < 3 * dm.Du; i += blockDim.x) Wsm[3 * dm.Di + i] = a.w_user[i];
    float* rows = sm + sp.warp_off + warp * sp.warp_stride;
    int32_t* kbuf = reinterpret_cast<int32_t*>(rows + sp.kbuf_off);
    float* dots = rows + sp.dots_off;
    float* wts = rows + sp.wts_off;      // w1[K] | w2[K] | dz1[K] | dz2[K]   (strides of 32)
    float* dbuf = rows + sp.dbuf_off;    // d user_side [Ds] | d item_side [Ds] | d atten_info [4K]
    float* acc = rows + sp.acc_off;      // dW1_item [Di] | dW2_item [Di] | dW1_user [Du] | dW2_user [Du]
    const SliceGeom g = slice_geom(dm);
    const int K = dm.K, Di = dm.Di, Du = dm.Du, Ds = dm.Ds, d = dm.d;
    const int nacc = 2 * Di + 2 * Du;
    const float invK = 1.0f / (float)K;
    for (int c = lane; c < nacc; c += 32) acc[c] = 0.f;
    __syncthreads();
    const int64_t M = (int64_t)dm.B * dm.T;
    for (int64_t slice = (int64_t)blockIdx.x * warps + warp; slice < M; slice += (int64_t)gridDim.x * warps) {
        const int b = (int)(slice / dm.T), t = (int)(slice - (int64_t)b * dm.T);
        if (t >= a.length[b]) {
            if (lane == 0) { a.sdz[slice * 2] = 0.f; a.sdz[slice * 2 + 1] = 0.f; }
            continue;   // positions of dead slices carry key 0: their gradient rows are never read
        }
        stage_slice(rows, kbuf, dm, g, a.emb, a.keys, slice, lane);
        {   // incoming gradients of this slice (overlaps the row gather)
            const float* dxu = a.dxu + slice * Ds; const float* dxi = a.dxi + slice * Ds;
            const float* dinfo = a.dkey + slice * a.ldkey + a.key_off;
            for (int c = lane; c < Ds; c += 32) { dbuf[c] = dxu[c]; dbuf[Ds + c] = dxi[c]; }
            for (int c = lane; c < 4 * K; c += 32) dbuf[2 * Ds + c] = dinfo[c];
        }
        float w1 = 0.f, w2 = 0.f, r1 = 0.f, r2 = 0.f;
        if (lane < K) {
            const float* sr = a.save_r + slice * 2 * K; const float* sw = a.save_w + slice * 2 * K;
            r1 = sr[lane]; r2 = sr[K + lane]; w1 = sw[lane]; w2 = sw[K + lane];
        }
        cp_async_wait<0>();
        __syncwarp();
        // (1) dw: dot of every seq1 row (seg0, seg2) with its field's slice of dout1
        for (int r = lane; r < g.nrows; r += 32) {
            int seg, i, field;
            row_decode(g, dm, r, seg, i, field);
            if (seg == 0 || seg == 2) {
                const float4* dv = reinterpret_cast<const float4*>(dbuf + (seg == 0 ? 0 : Di) + field * d);
                float s = 0.f;
                for (int c4 = 0; c4 < g.cpr; ++c4)
                    s += dot4(*reinterpret_cast<const float4*>(rows + chunk_off(g, r, c4)), dv[c4]);
                dots[r] = s;
            }
        }
        __syncwarp();
        // (2) per-neighbor scalars in lanes
        float dw1 = 0.f, dw2 = 0.f;
        if (lane < K) {
            for (int f = 0; f < dm.fi; ++f) dw1 += dots[lane * dm.fi + f];
            for (int f = 0; f < dm.fu; ++f) dw2 += dots[2 * g.nfi + lane * dm.fu + f];
        }
        const float* dinf = dbuf + 2 * Ds;
        const float dot1 = warp_sum(w1 * dw1), dot2 = warp_sum(w2 * dw2);
        const float tail1 = warp_sum(lane < K ? dinf[K + lane] : 0.f), tail2 = warp_sum(lane < K ? dinf[3 * K + lane] : 0.f);
        float dz1 = 0.f, dz2 = 0.f;
        if (lane < K) {
            const float dr1 = (float)K * dinf[lane] + tail1 + w1 * (dw1 - dot1);
            const float dr2 = (float)K * dinf[2 * K + lane] + tail2 + w2 * (dw2 - dot2);
            dz1 = r1 > 0.f ? dr1 : 0.f;
            dz2 = r2 > 0.f ? dr2 : 0.f;
            wts[lane] = w1; wts[32 + lane] = w2; wts[64 + lane] = dz1; wts[96 + lane] = dz2;
        }
        const float sdz1 = warp_sum(dz1), sdz2 = warp_sum(dz2);
        if (lane == 0) { a.sdz[slice * 2] = sdz1; a.sdz[slice * 2 + 1] = sdz2; }
        __syncwarp();
        // (3) per-position gradient rows, one 16-byte chunk per lane, written in position order (coalesced)
        {
            const int total = g.nrows * g.cpr;
            const int shift = 31 - __clz(g.cpr);
            float* gr = a.grad_rows;
            float* gb0 = gr + (dm.off_u1 + slice * g.nfi) * d;
            float* gb1 = gr + (dm.off_i2 + slice * g.nfi) * d;
            float* gb2 = gr + (dm.off_u2 + slice * g.nfu) * d;
            float* gb3 = gr + (dm.off_i1 + slice * g.nfu) * d;
            for (int q = lane; q < total; q += 32) {
                const int r = q >> shift, c4 = q & (g.cpr - 1);
                int seg, i, field;
                row_decode(g, dm, r, seg, i, field);
                const int col = field * d + c4 * 4;
                float a_i, dz; const float* dv; const float* wv; float* dst;
                if (seg == 0) { a_i = wts[i]; dz = wts[64 + i]; dv = dbuf + col; wv = Wsm + Di + col; dst = gb0 + (int64_t)r * d; }
                else if (seg == 1) { a_i = invK; dz = wts[64 + i]; dv = dbuf + Ds + Du + col; wv = Wsm + 2 * Di + col; dst = gb1 + (int64_t)(r - g.nfi) * d; }
                else if (seg == 2) { a_i = wts[32 + i]; dz = wts[96 + i]; dv = dbuf + Di + col; wv = Wsm + 3 * Di + Du + col; dst = gb2 + (int64_t)(r - 2 * g.nfi) * d; }
                else { a_i = invK; dz = wts[96 + i]; dv = dbuf + Ds + col; wv = Wsm + 3 * Di + 2 * Du + col; dst = gb3 + (int64_t)(r - 2 * g.nfi - g.nfu) * d; }
                const float4 dvv = *reinterpret_cast<const float4*>(dv);
                const float4 wvv = *reinterpret_cast<const float4*>(wv);
                float4 o;
                o.x = a_i * dvv.x + dz * wvv.x; o.y = a_i * dvv.y + dz * wvv.y;
                o.z = a_i * dvv.z + dz * wvv.z; o.w = a_i * dvv.w + dz * wvv.w;
                *reinterpret_cast<float4*>(dst + c4 * 4) = o;
            }
        }
        // (4) co-attention kernel gradient: dW1 += sum_i dz_i seq1[i], dW2 += sum_i dz_i seq2[i] (per-warp accumulators)
        {
            const int nchunk = nacc >> 2;
            for (int e4 = lane; e4 < nchunk; e4 += 32) {
                const int e = e4 << 2;
                int segbase, f, c, zsel;
                if (e < Di) { segbase = 0; f = dm.fi; c = e; zsel = 64; }
                else if (e < 2 * Di) { segbase = g.nfi; f = dm.fi; c = e - Di; zsel = 64; }
                else if (e < 2 * Di + Du) { segbase = 2 * g.nfi; f = dm.fu; c = e - 2 * Di; zsel = 96; }
                else { segbase = 2 * g.nfi + g.nfu; f = dm.fu; c = e - 2 * Di - Du; zsel = 96; }
                const int field = c / d, c4 = (c - field * d) >> 2;
                float4 s = *reinterpret_cast<float4*>(acc + e);
                for (int i = 0; i < K; ++i) {
                    const float4 v = *reinterpret_cast<const float4*>(rows + chunk_off(g, segbase + i * f + field, c4));
                    const float z = wts[zsel + i];
                    s.x = fmaf(z, v.x, s.x); s.y = fmaf(z, v.y, s.y); s.z = fmaf(z, v.z, s.z); s.w = fmaf(z, v.w, s.w);
                }
                *reinterpret_cast<float4*>(acc + e) = s;
            }
        }
        __syncwarp();
    }
    __syncthreads();
    // fixed-order sum over the CTA's warps -> one partial row per CTA
    for (int c = threadIdx.x; c < nacc; c += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < warps; ++w) s += sm[sp.warp_off + w * sp.warp_stride + sp.acc_off + c];
        a.partials[(int64_t)blockIdx.x * nacc + c] = s;
    }
}

int coatt_bwd_num_ctas() { return num_sms() * 8; }

void launch_coatt_bwd(cudaStream_t st, const Dims& dm, const CoattBwdArgs& a) {
    const int warps = 4;
    size_t smem;
    CoattSmem sp = coatt_plan(dm, warps, true, &smem);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaFuncSetAttribute(coatt_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = smem;
    }
    coatt_bwd_kernel<<<a.n_partials, warps * 32, smem, st>>>(dm, a, sp);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------
// target rows backward: one warp per sample, grid-stride, per-CTA partials for dWt / dbias.
__global__ void target_bwd_kernel(Dims dm, TargetBwdArgs a) {
    extern __shared__ float sm[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nacc = dm.Di + 1 + dm.Du + 1;
    float* acc = sm + warp * nacc;
    for (int c = lane; c < nacc; c += 32) acc[c] = 0.f;
    __syncwarp();
    const bool coatt = a.w_item != nullptr;
    for (int b = blockIdx.x * warps + warp; b < dm.B; b += gridDim.x * warps) {
        float s_item = 0.f, s_user = 0.f;
        if (coatt) {
            const int len = min(a.length[b], dm.T);
            for (int t = 0; t < len; ++t) {   // fixed order
                s_item += a.sdz[((int64_t)b * dm.T + t) * 2];
                s_user += a.sdz[((int64_t)b * dm.T + t) * 2 + 1];
            }
        }
        const float* q0 = a.q0 + (int64_t)b * dm.Ds;
        const float* dq0 = a.dq0 ? a.dq0 + (int64_t)b * dm.Ds : nullptr;
        const float* dfc = a.dfc_in + (int64_t)b * a.ldfc + a.fc_off;   // [d target_item (Di) | d target_user (Du)]
        float* g_tu = a.grad_rows + (dm.off_tu + (int64_t)b * dm.fu) * dm.d;
        float* g_ti = a.grad_rows + (dm.off_ti + (int64_t)b * dm.fi) * dm.d;
        for (int c = lane; c < dm.Du; c += 32) {
            float g = dfc[dm.Di + c] + (dq0 ? dq0[c] : 0.f);
            if (coatt) { g += s_user * a.w_user[c]; acc[dm.Di + 1 + c] += s_user * q0[c]; }
            g_tu[c] = g;
        }
        for (int c = lane; c < dm.Di; c += 32) {
            float g = dfc[c] + (dq0 ? dq0[dm.Du + c] : 0.f);
            if (coatt) { g += s_item * a.w_item[c]; acc[c] += s_item * q0[dm.Du + c]; }
            g_ti[c] = g;
        }
        if (lane == 0) { acc[dm.Di] += s_item; acc[dm.Di + 1 + dm.Du] += s_user; }
        __syncwarp();
    }
    __syncthreads();
    for (int c = threadIdx.x; c < nacc; c += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < warps; ++w) s += sm[w * nacc + c];
        a.partials[(int64_t)blockIdx.x * nacc + c] = s;
    }
}

int target_bwd_num_ctas() { return num_sms(); }

void launch_target_bwd(cudaStream_t st, const Dims& dm, const TargetBwdArgs& a) {
    const int warps = 4;
    size_t smem = (size_t)warps * (dm.Di + dm.Du + 2) * sizeof(float);
    target_bwd_kernel<<<a.n_partials, warps * 32, smem, st>>>(dm, a);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------
// co-attention kernel gradient [Wt | W1 | W2] + bias, for both co-attentions: one warp per output element,
// lanes stride over the per-CTA partial rows, fixed-shape butterfly at the end (deterministic).
__global__ void coatt_grad_reduce_kernel(Dims dm, const float* __restrict__ cp, int n_coatt,
                                         const float* __restrict__ tp, int n_target,
                                         float* g_w_item, float* g_b_item, float* g_w_user, float* g_b_user) {
    const int nacc_c = 2 * dm.Di + 2 * dm.Du;
    const int nacc_t = dm.Di + 1 + dm.Du + 1;
    const int total = 3 * dm.Di + 1 + 3 * dm.Du + 1;
    const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= total) return;
    // map e -> (source buffer, column, destination)
    const float* src; int col, n, stride; float* dst;
    if (e < 3 * dm.Di) {
        if (e < dm.Di) { src = tp; col = e; n = n_target; stride = nacc_t; }
        else { src = cp; col = e - dm.Di; n = n_coatt; stride = nacc_c; }
        dst = g_w_item + e;
    } else if (e == 3 * dm.Di) {
        src = tp; col = dm.Di; n = n_target; stride = nacc_t; dst = g_b_item;
    } else if (e < 3 * dm.Di + 1 + 3 * dm.Du) {
        int k = e - (3 * dm.Di + 1);
        if (k < dm.Du) { src = tp; col = dm.Di + 1 + k; n = n_target; stride = nacc_t; }
        else { src = cp; col = 2 * dm.Di + (k - dm.Du); n = n_coatt; stride = nacc_c; }
        dst = g_w_user + k;
    } else {
        src = tp; col = dm.Di + 1 + dm.Du; n = n_target; stride = nacc_t; dst = g_b_user;
    }
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s += src[(int64_t)i * stride + col];
    s = warp_sum(s);
    if (lane == 0) *dst = s;
}

void launch_coatt_grad_reduce(cudaStream_t st, const Dims& dm, const float* coatt_partials, int n_coatt,
                              const float* target_partials, int n_target,
                              float* g_w_item, float* g_b_item, float* g_w_user, float* g_b_user) {
    int total = 3 * dm.Di + 1 + 3 * dm.Du + 1;
    coatt_grad_reduce_kernel<<<(total * 32 + 127) / 128, 128, 0, st>>>(dm, coatt_partials, n_coatt, target_partials,
                                                                       n_target, g_w_item, g_b_item, g_w_user, g_b_user);
    ++g_launch_count;
}

}  // namespace score
