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
//
// Position space.  Every id of a batch owns one "position" p in [0, N): the sanitized id lives at keys[p], its
// embedding-gradient row at grad_rows[p*d ..].  Positions are SLICE-MAJOR: the NROWS = K*(2*if + 2*uf) ids of the
// (b,t) slice s sit at [s*NROWS, (s+1)*NROWS) in the order
//   seg0 = user_1hop (K*if) | seg1 = item_2hop (K*if) | seg2 = user_2hop (K*uf) | seg3 = item_1hop (K*uf)
// (co-attention #1 = (seg0, seg1, target_item), #2 = (seg2, seg3, target_user), score.py:196-197), followed by
// target_user [B*uf] and target_item [B*if].  One warp owns one slice: its ids are one contiguous run, its table rows
// land in shared memory in the same order, and its gradient rows leave as one contiguous block.
//
// The kernels are instruction-bound before they are bandwidth-bound (ncu, profiles/: 1 650 warp instructions per
// Taobao slice with run-time geometry), so they are compiled once per geometry (K, if, uf, d) of the reference's data
// sets - every index computation folds to constants, every loop unrolls - with a run-time-geometry instance as the
// fallback for any other configuration.
#include "kernels.h"

namespace score {

// ------------------------------------------------------------------------------------------
// keys[p] = sanitized id of position p: 0 for masked slices (t >= length[b]) and ids that are out of range (flagged).
// The six id tensors are read where they lie (device batches: the caller's tensors, no staging copy) through a pointer
// table in device memory, so a captured graph stays valid from batch to batch.  Four positions per thread.  In LAZY
// optimizer mode the same pass claims the stale rows among the keys (atomic exchange on last_step, plain pre-check
// first) and appends (row, last step) to the compact list emb_replay_kernel works through (scatter.cu).
// Work list of the lean co-attention kernels: the live slices (t < length[b]) as (b << 8) | t in (b, t) order, their
// number at live[B*T].  CTA `cta` of `ncta` covers 256 samples: the slices before its first sample are counted by the
// whole CTA (<= B loads), its own 256 lengths are scanned, every thread writes its sample's entries.  No atomics: the
// list order is fixed, so the per-warp accumulation order of the backward kernel is reproducible.
__device__ void build_live_list(const Dims& dm, const int32_t* __restrict__ length, int32_t* __restrict__ live, int cta, int ncta) {
    __shared__ int red[8], wtot[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = cta * 256;
    int s = 0;
    for (int b = tid; b < b0; b += 256) s += min(max(length[b], 0), dm.T);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
    const int b = b0 + tid;
    const int len = b < dm.B ? min(max(length[b], 0), dm.T) : 0;
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(FULL_MASK, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 0) red[warp] = s;
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    int off = incl - len;
#pragma unroll
    for (int w = 0; w < 8; ++w) { off += red[w]; if (w < warp) off += wtot[w]; }
    for (int t = 0; t < len; ++t) live[off + t] = (b << 8) | t;
    if (cta == ncta - 1 && tid == 255) live[(int64_t)dm.B * dm.T] = off + len;
}

__global__ void build_keys_kernel(Dims dm, const BatchPtrs* __restrict__ bpp, int32_t* __restrict__ keys,
                                  int32_t* __restrict__ label_out, int32_t* __restrict__ length_out,
                                  int32_t* __restrict__ err_flag, ClaimArgs ca, int32_t* __restrict__ live, int n_live_ctas) {
    pdl_enter();
    const BatchPtrs bp = *bpp;
    if (blockIdx.x >= gridDim.x - n_live_ctas) {   // the last CTAs build the live-slice list (whole CTA, uniform)
        build_live_list(dm, bp.length, live, blockIdx.x - (gridDim.x - n_live_ctas), n_live_ctas);
        return;
    }
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gtid < dm.B) {
        const int32_t lb = bp.label[gtid], ln = bp.length[gtid];
        label_out[gtid] = lb; length_out[gtid] = ln;
    }
    const int64_t p0 = gtid * 4;
    const int lane = threadIdx.x & 31;
    const uint32_t nrows = (uint32_t)dm.nrows, nfi = (uint32_t)(dm.K * dm.fi), nfu = (uint32_t)(dm.K * dm.fu);
    int32_t id[4] = {0, 0, 0, 0};
    bool bad = false;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int64_t p = p0 + u;
        int32_t v = 0;
        if (p < dm.off_tu) {
            const uint32_t slice = (uint32_t)p / nrows, r = (uint32_t)p - slice * nrows;
            const uint32_t b = slice / (uint32_t)dm.T, t = slice - b * (uint32_t)dm.T;
            if ((int32_t)t < bp.length[b]) {
                // hop1_only (RRN, slice_model.py:155-157): user_2hop / item_2hop are fed but have no consumer -> key 0
                if (r < nfi) v = bp.u1[(int64_t)slice * nfi + r];
                else if (r < 2 * nfi) v = dm.hop1_only ? 0 : bp.i2[(int64_t)slice * nfi + (r - nfi)];
                else if (r < 2 * nfi + nfu) v = dm.hop1_only ? 0 : bp.u2[(int64_t)slice * nfu + (r - 2 * nfi)];
                else v = bp.i1[(int64_t)slice * nfu + (r - 2 * nfi - nfu)];
            }
        } else if (p < dm.off_ti) v = bp.tu[p - dm.off_tu];
        else if (p < dm.N) v = bp.ti[p - dm.off_ti];
        if (v < 0 || (int64_t)v >= dm.V) { bad = true; v = 0; }
        id[u] = v;
    }
    if (bad) atomicExch(err_flag, 1);
    if (p0 + 3 < dm.N) *reinterpret_cast<int4*>(keys + p0) = make_int4(id[0], id[1], id[2], id[3]);
    else {
#pragma unroll
        for (int u = 0; u < 4; ++u) if (p0 + u < dm.N) keys[p0 + u] = id[u];
    }
    if (!ca.last_step) return;   // uniform
    int32_t* counter = ca.counter;
    if (ca.pingpong) {
        const int sel = ca.hp->seq & 1;
        counter += sel;
        if (gtid == 0) ca.counter[sel ^ 1] = 0;   // the next launch's counter (idle now: its reader finished a launch ago)
    }
    const int upto = ca.hp->step - 1;
    int seen[4], old[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) seen[u] = id[u] != 0 ? ca.last_step[id[u]] : upto;   // four independent loads
    int nwin = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        old[u] = upto;
        // unsigned compare: last_step == -1 marks a row no optimizer step has touched (m = v = 0: every skipped
        // zero-gradient step is exactly a no-op) - it is current by construction and never claimed
        if ((uint32_t)seen[u] < (uint32_t)upto) old[u] = atomicExch(&ca.last_step[id[u]], upto);
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
    if (lane == 31) base = atomicAdd(counter, total);
    base = __shfl_sync(FULL_MASK, base, 31);
    int idx = base + incl - nwin;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (old[u] < upto) { ca.list[2 * idx] = id[u]; ca.list[2 * idx + 1] = old[u]; ++idx; }
}

void launch_build_keys(cudaStream_t st, const Dims& dm, const BatchPtrs* bp_dev, int32_t* keys, int32_t* label_out,
                       int32_t* length_out, int32_t* err_flag, const ClaimArgs* claim, int32_t* live) {
    ClaimArgs ca{};
    if (claim) {
        ca = *claim;
        if (!ca.pingpong) cudaMemsetAsync(ca.counter, 0, sizeof(int32_t), st);
    }
    const int threads = 256;
    int64_t work = (dm.N + 3) / 4;
    if (work < dm.B) work = dm.B;
    const int64_t blocks = (work + threads - 1) / threads;
    const int n_live = (live && dm.T <= 255) ? (dm.B + 255) / 256 : 0;
    launch_chain(build_keys_kernel, dim3((unsigned)(blocks + n_live)), dim3(threads), 0, st, dm, bp_dev, keys, label_out, length_out, err_flag, ca,
                 live, n_live);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------
// stage `nrows` table rows (ids at key_ptr[0..nrows)) into smem dst[nrows][d]; whole warp cooperates, consecutive
// lanes fetch consecutive 16-byte chunks (a d=16 row is 4 lanes, 64 B contiguous).
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ emb, int es, const int32_t* key_ptr,
                                           int nrows, int d, int lane) {
    const int cpr = d >> 2;   // 16-byte chunks per row
    const int total = nrows * cpr;
    for (int q = lane; q < total; q += 32) {
        int r = q / cpr, c = q - r * cpr;
        int32_t id = key_ptr[r];
        const float* src = emb + (int64_t)id * es + c * 4;
        cp_async16(dst + r * d + c * 4, id != 0 ? src : emb, id != 0 ? 16 : 0);
    }
}

// ------------------------------------------------------------------------------------------
// target rows: one warp per sample b
__global__ void target_fwd_kernel(Dims dm, TargetArgs a) {
    pdl_enter();
    extern __shared__ float sm[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* buf = sm + warp * dm.Ds;   // [tu (Du) | ti (Di)]
    int b = blockIdx.x * warps + warp;
    if (b >= dm.B) return;
    stage_rows(buf, a.emb, a.es, a.keys + dm.off_tu + (int64_t)b * dm.fu, dm.fu, dm.d, lane);
    stage_rows(buf + dm.Du, a.emb, a.es, a.keys + dm.off_ti + (int64_t)b * dm.fi, dm.fi, dm.d, lane);
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
    launch_chain(target_fwd_kernel, dim3((dm.B + warps - 1) / warps), dim3(warps * 32), smem, st, dm, a);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------
// Slice geometry: compile-time for the reference's data sets, run-time otherwise.  Same accessors, so the kernels
// below are written once.
template <int K_, int FI_, int FU_, int D_>
struct GeomS {
    static constexpr bool kStatic = true;
    __device__ __forceinline__ explicit GeomS(const Dims&) {}
    __device__ __forceinline__ constexpr int K() const { return K_; }
    __device__ __forceinline__ constexpr int FI() const { return FI_; }
    __device__ __forceinline__ constexpr int FU() const { return FU_; }
    __device__ __forceinline__ constexpr int D() const { return D_; }
    __device__ __forceinline__ constexpr int cpr() const { return D_ / 4; }
    __device__ __forceinline__ constexpr int logcpr() const {
        return D_ == 4 ? 0 : D_ == 8 ? 1 : D_ == 16 ? 2 : D_ == 32 ? 3 : D_ == 64 ? 4 : 5;
    }
};
struct GeomD {
    static constexpr bool kStatic = false;
    int k, fi, fu, d, lc;
    __device__ __forceinline__ explicit GeomD(const Dims& dm) : k(dm.K), fi(dm.fi), fu(dm.fu), d(dm.d) { lc = 31 - __clz(dm.d >> 2); }
    __device__ __forceinline__ int K() const { return k; }
    __device__ __forceinline__ int FI() const { return fi; }
    __device__ __forceinline__ int FU() const { return fu; }
    __device__ __forceinline__ int D() const { return d; }
    __device__ __forceinline__ int cpr() const { return d >> 2; }
    __device__ __forceinline__ int logcpr() const { return lc; }
};
// first row, number of rows and fields per node of segment `seg`; float offset of the segment's slice of the
// co-attention kernels in the CTA's weight buffer [Wt_item | W1_item | W2_item | Wt_user | W1_user | W2_user]
template <class G>
__device__ __forceinline__ void seg_geom(const G& g, int seg, int& row0, int& nrow, int& F, int& wofs) {
    const int nfi = g.K() * g.FI(), nfu = g.K() * g.FU(), Di = g.FI() * g.D(), Du = g.FU() * g.D();
    row0 = seg == 0 ? 0 : seg == 1 ? nfi : seg == 2 ? 2 * nfi : 2 * nfi + nfu;
    nrow = seg < 2 ? nfi : nfu;
    F = seg < 2 ? g.FI() : g.FU();
    wofs = seg == 0 ? Di : seg == 1 ? 2 * Di : seg == 2 ? 3 * Di + Du : 3 * Di + 2 * Du;
}

// Shared memory per warp: the slice's table rows in position order (linear 16-byte chunks: chunk q of the slice at
// float offset 4q, so lane-per-chunk reads are conflict free), one dot product per row, the per-neighbor weights of
// both co-attentions, and (backward) the slice's incoming gradients and the warp's accumulators for the
// co-attention kernel gradient.
struct CoattSmem {
    int w_off;        // [3*Di + 3*Du] co-attention kernels (CTA-shared)
    int warp_off;     // per-warp region start (floats)
    int warp_stride;  // per-warp floats
    int nrows;        // K * (2*fi + 2*fu)
    // per-warp layout (floats): rows[nrows*d] | dots[nrows] | wts[4*WS] | (bwd) dbuf[2*Ds + 4K] | (bwd) acc[2*Di+2*Du]
    int dots_off, wts_off, dbuf_off, acc_off;
};
__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }
// stride between the per-neighbor weight arrays of a warp: 8 banks apart, so lanes that read entry i of different
// arrays (pooling: one array per segment) do not collide
constexpr int WS = 40;

__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// Gather one (b,t) slice: every lane fetches its 16-byte chunks with cp.async; the id of a row is read by the lanes
// that need it (one broadcast load per row, all of them independent), never staged first.
template <class G>
__device__ __forceinline__ void stage_slice(const G& g, float* rows, const float* __restrict__ emb, int es,
                                            const int32_t* __restrict__ ks, int nrows, int lane) {
    const int total = nrows << g.logcpr();
#pragma unroll
    for (int q0 = 0; q0 < total; q0 += 32) {
        const int q = q0 + lane;
        if (q < total) {
            const int r = q >> g.logcpr(), c4 = q & (g.cpr() - 1);
            const int32_t id = __ldg(ks + r);
            const float* src = emb + (int64_t)id * es + c4 * 4;
            cp_async16(rows + q * 4, id != 0 ? src : emb, id != 0 ? 16 : 0);
        }
    }
    cp_async_commit();
}

// dots[row] = <row, vec[(row's field) * d ..]> for every row of segment `seg`: lane per 16-byte chunk, then a
// shuffle tree over the chunks of a row
template <class G>
__device__ __forceinline__ void seg_dots(const G& g, int seg, const float* rows, const float* vec, float* dots, int lane) {
    int row0, nrow, F, wofs;
    seg_geom(g, seg, row0, nrow, F, wofs);
    const int nch = nrow << g.logcpr(), fmod = F << g.logcpr();
    const float4* r4 = reinterpret_cast<const float4*>(rows) + (row0 << g.logcpr());
    const float4* v4 = reinterpret_cast<const float4*>(vec);
#pragma unroll
    for (int q0 = 0; q0 < nch; q0 += 32) {
        const int q = q0 + lane;
        float p = 0.f;
        if (q < nch) p = dot4(r4[q], v4[q % fmod]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            if (o < g.cpr()) p += __shfl_xor_sync(FULL_MASK, p, o);
        if (q < nch && (q & (g.cpr() - 1)) == 0) dots[row0 + (q >> g.logcpr())] = p;
    }
}

// reductions over the 16-lane half of a warp (both co-attentions side by side when K <= 16)
__device__ __forceinline__ float half_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float half_max(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

template <class G>
__global__ void __launch_bounds__(128) coatt_fwd_kernel(Dims dm, CoattArgs a, CoattSmem sp) {
    pdl_enter();
    extern __shared__ __align__(16) float sm[];
    const G g(dm);
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = g.K(), D = g.D(), Di = g.FI() * D, Du = g.FU() * D, Ds = Di + Du;
    const int nfi = K * g.FI(), nfu = K * g.FU(), nrows = 2 * nfi + 2 * nfu;
    float* Wsm = sm + sp.w_off;
    const bool sum_pool = a.sum_pool != 0;
    if (!sum_pool) {
        for (int i = threadIdx.x; i < 3 * Di; i += blockDim.x) Wsm[i] = a.w_item[i];
        for (int i = threadIdx.x; i < 3 * Du; i += blockDim.x) Wsm[3 * Di + i] = a.w_user[i];
    }
    float* rows = sm + sp.warp_off + warp * sp.warp_stride;
    float* dots = rows + sp.dots_off;
    float* wts = rows + sp.wts_off;      // w1[32] | w2[32] | weight of the seq2 neighbors [32]
    const float invK = 1.0f / (float)K;
    wts[lane] = 1.0f; wts[WS + lane] = 1.0f;          // sum_pool: every neighbor counts once
    wts[2 * WS + lane] = sum_pool ? 1.0f : invK;
    __syncthreads();
    const int M = dm.B * dm.T;
    for (int slice = blockIdx.x * warps + warp; slice < M; slice += gridDim.x * warps) {
        const int b = slice / dm.T, t = slice - b * dm.T;
        float* xu_g = a.xhg_u + (int64_t)slice * dm.ldxs[0]; float* xu_c = a.xhc_u + (int64_t)slice * dm.ldxs[0];
        float* xi_g = a.xhg_i + (int64_t)slice * dm.ldxs[1]; float* xi_c = a.xhc_i + (int64_t)slice * dm.ldxs[1];
        float* info = a.key + (int64_t)slice * a.ldkey + a.key_off;
        if (t >= a.length[b]) {   // dead slice: nothing downstream reads it, keep buffers finite
            for (int c = lane; c < dm.Dx[0]; c += 32) { xu_g[c] = 0.f; xu_c[c] = 0.f; }
            for (int c = lane; c < dm.Dx[1]; c += 32) { xi_g[c] = 0.f; xi_c[c] = 0.f; }
            if (!sum_pool) for (int c = lane; c < 4 * K; c += 32) info[c] = 0.f;
            continue;
        }
        stage_slice(g, rows, a.emb, a.es, a.keys + (int64_t)slice * nrows, nrows, lane);
        const float cz1 = sum_pool ? 0.f : a.c_item[b], cz2 = sum_pool ? 0.f : a.c_user[b];
        cp_async_wait<0>();
        __syncwarp();
        if (!sum_pool) {
        // (1) one dot product per row with its slice of the co-attention kernel (W1 for seq1 rows, W2 for seq2 rows)
#pragma unroll
        for (int seg = 0; seg < 4; ++seg) {
            int row0, nrow, F, wofs;
            seg_geom(g, seg, row0, nrow, F, wofs);
            seg_dots(g, seg, rows, Wsm + wofs, dots, lane);
        }
        __syncwarp();
        // (2) relatedness r_i = relu(target part + seq1[i] part + seq2[i] part), softmax over the K neighbors
        float* sr = a.save_r + (int64_t)slice * 2 * K; float* sw = a.save_w + (int64_t)slice * 2 * K;
        if (K <= 16) {   // lanes 0..15: co-attention #1, lanes 16..31: co-attention #2
            const int half = lane >> 4, i = lane & 15;
            const bool act = i < K;
            const int F = half ? g.FU() : g.FI();
            const int base = half ? 2 * nfi : 0, nf = half ? nfu : nfi;
            float z = half ? cz2 : cz1;
            if (act) {
                const int fmax = g.FI() > g.FU() ? g.FI() : g.FU();
#pragma unroll
                for (int f = 0; f < fmax; ++f)
                    if (f < F) z += dots[base + i * F + f] + dots[base + nf + i * F + f];
            }
            const float r = act ? fmaxf(z, 0.f) : 0.f;
            const float mx = half_max(act ? r : -INFINITY);
            const float e = act ? expf(r - mx) : 0.f;
            const float w = e / half_sum(e);
            const float s = half_sum(r);
            if (act) {
                wts[half * WS + i] = w;
                info[half * 2 * K + i] = (float)K * r;     // atten_info (score.py:165-166)
                info[half * 2 * K + K + i] = s;
                sr[half * K + i] = r; sw[half * K + i] = w;
            }
        } else {
            float z1 = cz1, z2 = cz2;
            if (lane < K) {
                for (int f = 0; f < g.FI(); ++f) z1 += dots[lane * g.FI() + f] + dots[nfi + lane * g.FI() + f];
                for (int f = 0; f < g.FU(); ++f) z2 += dots[2 * nfi + lane * g.FU() + f] + dots[2 * nfi + nfu + lane * g.FU() + f];
            }
            const bool act = lane < K;
            const float r1 = act ? fmaxf(z1, 0.f) : 0.f, r2 = act ? fmaxf(z2, 0.f) : 0.f;
            const float m1 = warp_max(act ? r1 : -INFINITY), m2 = warp_max(act ? r2 : -INFINITY);
            const float e1 = act ? expf(r1 - m1) : 0.f, e2 = act ? expf(r2 - m2) : 0.f;
            const float w1 = e1 / warp_sum(e1), w2 = e2 / warp_sum(e2);
            const float s1 = warp_sum(r1), s2 = warp_sum(r2);
            if (act) {
                wts[lane] = w1; wts[WS + lane] = w2;
                info[lane] = (float)K * r1; info[K + lane] = s1;
                info[2 * K + lane] = (float)K * r2; info[3 * K + lane] = s2;
                sr[lane] = r1; sr[K + lane] = r2; sw[lane] = w1; sw[K + lane] = w2;
            }
        }
        }   // !sum_pool
        __syncwarp();
        // (3) pooling, one 16-byte output chunk per lane:
        //   user_side = [sum_i w1_i user_1hop[i] | sum_i w2_i user_2hop[i]],  item_side = [mean item_1hop | mean item_2hop]
        const int nchunk = (2 * Ds) >> 2;
        const float4* r4 = reinterpret_cast<const float4*>(rows);
#pragma unroll
        for (int e0 = 0; e0 < nchunk; e0 += 32) {
            const int e4 = e0 + lane;
            if (e4 < nchunk) {
                const int e = e4 << 2;
                // RRN: user side = sum of user_1hop, item side = sum of item_1hop; the 2-hop halves do not exist
                if (dm.hop1_only && ((e >= Di && e < Ds) || e >= Ds + Du)) continue;
                int row0, F, c, wsel; float* dst0; float* dst1;
                if (e < Di) { row0 = 0; F = g.FI(); c = e; wsel = 0; dst0 = xu_g + e; dst1 = xu_c + e; }
                else if (e < Ds) { row0 = 2 * nfi; F = g.FU(); c = e - Di; wsel = WS; dst0 = xu_g + e; dst1 = xu_c + e; }
                else if (e < Ds + Du) { row0 = 2 * nfi + nfu; F = g.FU(); c = e - Ds; wsel = 2 * WS; dst0 = xi_g + c; dst1 = xi_c + c; }
                else { row0 = nfi; F = g.FI(); c = e - Ds - Du; wsel = 2 * WS; dst0 = xi_g + Du + c; dst1 = xi_c + Du + c; }
                const float4* src = r4 + (row0 << g.logcpr()) + (c >> 2);   // chunk (field, c4) of neighbor 0
                const int stride = F << g.logcpr();
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    const float4 v = src[i * stride];
                    const float w = wts[wsel + i];
                    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
                }
                *reinterpret_cast<float4*>(dst0) = acc;
                *reinterpret_cast<float4*>(dst1) = acc;
            }
        }
        __syncwarp();
    }
}

static CoattSmem coatt_plan(const Dims& dm, int warps, bool bwd, size_t* bytes) {
    CoattSmem sp;
    sp.w_off = 0;
    sp.warp_off = round4(3 * dm.Di + 3 * dm.Du);
    sp.nrows = dm.K * (2 * dm.fi + 2 * dm.fu);
    sp.dots_off = sp.nrows * dm.d;
    sp.wts_off = sp.dots_off + round4(sp.nrows);
    sp.dbuf_off = sp.wts_off + 4 * WS;
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

template <class G>
static void coatt_fwd_launch(cudaStream_t st, const Dims& dm, const CoattArgs& a) {
    const int warps = 4;
    size_t smem;
    CoattSmem sp = coatt_plan(dm, warps, false, &smem);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaFuncSetAttribute(coatt_fwd_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = smem;
    }
    int64_t M = (int64_t)dm.B * dm.T;
    int64_t want = (M + warps - 1) / warps;
    int64_t cap = (int64_t)num_sms() * 32;
    int grid = (int)(want < cap ? want : cap);
    launch_chain(coatt_fwd_kernel<G>, dim3(grid), dim3(warps * 32), smem, st, dm, a, sp);
}

// geometries of the reference's data sets (train_score.py:23-54) + the large-vocab config; anything else: run-time
#define SCORE_GEOM_DISPATCH(FN, ...)                                                                         \
    do {                                                                                                     \
        if (dm.K == 10 && dm.fi == 2 && dm.fu == 1 && dm.d == 16) FN<GeomS<10, 2, 1, 16>>(__VA_ARGS__);      \
        else if (dm.K == 10 && dm.fi == 4 && dm.fu == 3 && dm.d == 16) FN<GeomS<10, 4, 3, 16>>(__VA_ARGS__); \
        else if (dm.K == 10 && dm.fi == 5 && dm.fu == 1 && dm.d == 16) FN<GeomS<10, 5, 1, 16>>(__VA_ARGS__); \
        else if (dm.K == 20 && dm.fi == 5 && dm.fu == 1 && dm.d == 16) FN<GeomS<20, 5, 1, 16>>(__VA_ARGS__); \
        else if (dm.K == 10 && dm.fi == 1 && dm.fu == 1 && dm.d == 64) FN<GeomS<10, 1, 1, 64>>(__VA_ARGS__); \
        else FN<GeomD>(__VA_ARGS__);                                                                         \
    } while (0)

void launch_coatt_fwd(cudaStream_t st, const Dims& dm, const CoattArgs& a) {
    ++g_launch_count;
    if (try_coatt_fwd_lean(st, dm, a, a.live)) return;
    SCORE_GEOM_DISPATCH(coatt_fwd_launch, st, dm, a);
}

// ------------------------------------------------------------------------------------------
// backward of both co-attentions of one slice (same staging as forward; the rows are re-gathered, mostly from L2,
// instead of being stored by the forward pass):
//   dw_i = dout1 . seq1[i]                    dr_i = K*dinfo[i] + sum_j dinfo[K+j] + w_i (dw_i - sum_k w_k dw_k)
//   dz_i = dr_i [r_i > 0]                     dseq1[i] = w_i dout1 + dz_i W1      dseq2[i] = dout2 / K + dz_i W2
//   dW1 += sum_i dz_i seq1[i]                 dW2 += sum_i dz_i seq2[i]           sdz = sum_i dz_i
template <class G>
__global__ void __launch_bounds__(128) coatt_bwd_kernel(Dims dm, CoattBwdArgs a, CoattSmem sp) {
    pdl_enter();
    extern __shared__ __align__(16) float sm[];
    const G g(dm);
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = g.K(), D = g.D(), Di = g.FI() * D, Du = g.FU() * D, Ds = Di + Du;
    const int nfi = K * g.FI(), nfu = K * g.FU(), nrows = 2 * nfi + 2 * nfu;
    float* Wsm = sm + sp.w_off;
    const bool sum_pool = a.sum_pool != 0;
    for (int i = threadIdx.x; i < 3 * Di; i += blockDim.x) Wsm[i] = sum_pool ? 0.f : a.w_item[i];
    for (int i = threadIdx.x; i < 3 * Du; i += blockDim.x) Wsm[3 * Di + i] = sum_pool ? 0.f : a.w_user[i];
    float* rows = sm + sp.warp_off + warp * sp.warp_stride;
    float* dots = rows + sp.dots_off;
    float* wts = rows + sp.wts_off;      // a1[32] (w1) | a2[32] (w2) | dz1[32] | dz2[32]
    wts[lane] = 1.0f; wts[WS + lane] = 1.0f; wts[2 * WS + lane] = 0.f; wts[3 * WS + lane] = 0.f;   // the sum_pool constants
    float* dbuf = rows + sp.dbuf_off;    // d user_side [Ds] | d item_side [Ds] | d atten_info [4K]
    float* acc = rows + sp.acc_off;      // dW1_item [Di] | dW2_item [Di] | dW1_user [Du] | dW2_user [Du]
    const int nacc = 2 * Di + 2 * Du;
    const float invK = sum_pool ? 1.0f : 1.0f / (float)K;   // weight of a seq2 neighbor in its pooled output
    for (int c = lane; c < nacc; c += 32) acc[c] = 0.f;
    __syncthreads();
    const int M = dm.B * dm.T;
    for (int slice = blockIdx.x * warps + warp; slice < M; slice += gridDim.x * warps) {
        const int b = slice / dm.T, t = slice - b * dm.T;
        if (t >= a.length[b]) {
            if (lane == 0) { a.sdz[(int64_t)slice * 2] = 0.f; a.sdz[(int64_t)slice * 2 + 1] = 0.f; }
            continue;   // positions of dead slices carry key 0: their gradient rows are never read
        }
        stage_slice(g, rows, a.emb, a.es, a.keys + (int64_t)slice * nrows, nrows, lane);
        {   // incoming gradients of this slice (overlaps the row gather)
            const float4* dxu = reinterpret_cast<const float4*>(a.dxu + (int64_t)slice * dm.Dx[0]);
            const float4* dxi = reinterpret_cast<const float4*>(a.dxi + (int64_t)slice * dm.Dx[1]);
            float4* d4 = reinterpret_cast<float4*>(dbuf);
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);   // RRN: the 2-hop halves have no gradient
            for (int c = lane; c < (Ds >> 2); c += 32) {
                d4[c] = c < (dm.Dx[0] >> 2) ? dxu[c] : z4;
                d4[(Ds >> 2) + c] = c < (dm.Dx[1] >> 2) ? dxi[c] : z4;
            }
            if (a.dkey) {
                const float* dinfo = a.dkey + (int64_t)slice * a.ldkey + a.key_off;
                for (int c = lane; c < 4 * K; c += 32) dbuf[2 * Ds + c] = dinfo[c];
            } else {
                for (int c = lane; c < 4 * K; c += 32) dbuf[2 * Ds + c] = 0.f;
            }
        }
        const float* sr = a.save_r + (int64_t)slice * 2 * K; const float* sw = a.save_w + (int64_t)slice * 2 * K;
        float w_h = 0.f, r_h = 0.f, w1 = 0.f, w2 = 0.f, r1 = 0.f, r2 = 0.f;
        if (sum_pool) {
        } else if (K <= 16) {
            const int half = lane >> 4, i = lane & 15;
            if (i < K) { w_h = sw[half * K + i]; r_h = sr[half * K + i]; }
        } else if (lane < K) {
            r1 = sr[lane]; r2 = sr[K + lane]; w1 = sw[lane]; w2 = sw[K + lane];
        }
        cp_async_wait<0>();
        __syncwarp();
        if (sum_pool) {
            if (lane == 0) { a.sdz[(int64_t)slice * 2] = 0.f; a.sdz[(int64_t)slice * 2 + 1] = 0.f; }
        } else {
        // (1) dw: dot of every seq1 row (seg0, seg2) with its field's slice of dout1
        seg_dots(g, 0, rows, dbuf, dots, lane);
        seg_dots(g, 2, rows, dbuf + Di, dots, lane);
        __syncwarp();
        // (2) per-neighbor scalars in lanes
        const float* dinf = dbuf + 2 * Ds;
        if (K <= 16) {
            const int half = lane >> 4, i = lane & 15;
            const bool act = i < K;
            const int F = half ? g.FU() : g.FI();
            const int base = half ? 2 * nfi : 0;
            float dw = 0.f;
            if (act) {
                const int fmax = g.FI() > g.FU() ? g.FI() : g.FU();
#pragma unroll
                for (int f = 0; f < fmax; ++f)
                    if (f < F) dw += dots[base + i * F + f];
            }
            const float* di = dinf + half * 2 * K;
            const float dotw = half_sum(w_h * dw);
            const float tail = half_sum(act ? di[K + i] : 0.f);
            float dz = 0.f;
            if (act) {
                const float dr = (float)K * di[i] + tail + w_h * (dw - dotw);
                dz = r_h > 0.f ? dr : 0.f;
                wts[half * WS + i] = w_h; wts[2 * WS + half * WS + i] = dz;
            }
            const float sdz = half_sum(dz);
            if (i == 0) a.sdz[(int64_t)slice * 2 + half] = sdz;
        } else {
            float dw1 = 0.f, dw2 = 0.f;
            if (lane < K) {
                for (int f = 0; f < g.FI(); ++f) dw1 += dots[lane * g.FI() + f];
                for (int f = 0; f < g.FU(); ++f) dw2 += dots[2 * nfi + lane * g.FU() + f];
            }
            const float dot1 = warp_sum(w1 * dw1), dot2 = warp_sum(w2 * dw2);
            const float tail1 = warp_sum(lane < K ? dinf[K + lane] : 0.f), tail2 = warp_sum(lane < K ? dinf[3 * K + lane] : 0.f);
            float dz1 = 0.f, dz2 = 0.f;
            if (lane < K) {
                const float dr1 = (float)K * dinf[lane] + tail1 + w1 * (dw1 - dot1);
                const float dr2 = (float)K * dinf[2 * K + lane] + tail2 + w2 * (dw2 - dot2);
                dz1 = r1 > 0.f ? dr1 : 0.f;
                dz2 = r2 > 0.f ? dr2 : 0.f;
                wts[lane] = w1; wts[WS + lane] = w2; wts[2 * WS + lane] = dz1; wts[3 * WS + lane] = dz2;
            }
            const float sdz1 = warp_sum(dz1), sdz2 = warp_sum(dz2);
            if (lane == 0) { a.sdz[(int64_t)slice * 2] = sdz1; a.sdz[(int64_t)slice * 2 + 1] = sdz2; }
        }
        }   // !sum_pool
        __syncwarp();
        // (3) per-position gradient rows, one 16-byte chunk per lane; the slice's rows are one contiguous block
        {
            float4* gout = reinterpret_cast<float4*>(a.grad_rows) + ((int64_t)slice * nrows << g.logcpr());
#pragma unroll
            for (int seg = 0; seg < 4; ++seg) {
                int row0, nrow, F, wofs;
                seg_geom(g, seg, row0, nrow, F, wofs);
                const int nch = nrow << g.logcpr(), fmod = F << g.logcpr();
                // d(pooled output) of this segment and its slice of the co-attention kernel
                const float4* dv4 = reinterpret_cast<const float4*>(dbuf + (seg == 0 ? 0 : seg == 1 ? Ds + Du : seg == 2 ? Di : Ds));
                const float4* wv4 = reinterpret_cast<const float4*>(Wsm + wofs);
                const float* ai = wts + (seg == 0 ? 0 : WS);        // softmax weights (seq1 segments)
                const float* dzi = wts + (seg < 2 ? 2 * WS : 3 * WS);
                float4* go = gout + (row0 << g.logcpr());
#pragma unroll
                for (int q0 = 0; q0 < nch; q0 += 32) {
                    const int q = q0 + lane;
                    if (q < nch) {
                        const int i = q / fmod, fc = q - i * fmod;
                        const float a_i = (seg == 0 || seg == 2) ? ai[i] : invK;
                        const float dz = dzi[i];
                        const float4 dvv = dv4[fc], wvv = wv4[fc];
                        float4 o;
                        o.x = a_i * dvv.x + dz * wvv.x; o.y = a_i * dvv.y + dz * wvv.y;
                        o.z = a_i * dvv.z + dz * wvv.z; o.w = a_i * dvv.w + dz * wvv.w;
                        go[q] = o;
                    }
                }
            }
        }
        // (4) co-attention kernel gradient: dW1 += sum_i dz_i seq1[i], dW2 += sum_i dz_i seq2[i] (per-warp accumulators)
        if (!sum_pool) {
            const int nchunk = nacc >> 2;
            const float4* r4 = reinterpret_cast<const float4*>(rows);
#pragma unroll
            for (int e0 = 0; e0 < nchunk; e0 += 32) {
                const int e4 = e0 + lane;
                if (e4 < nchunk) {
                    const int e = e4 << 2;
                    int row0, F, c, zsel;
                    if (e < Di) { row0 = 0; F = g.FI(); c = e; zsel = 2 * WS; }
                    else if (e < 2 * Di) { row0 = nfi; F = g.FI(); c = e - Di; zsel = 2 * WS; }
                    else if (e < 2 * Di + Du) { row0 = 2 * nfi; F = g.FU(); c = e - 2 * Di; zsel = 3 * WS; }
                    else { row0 = 2 * nfi + nfu; F = g.FU(); c = e - 2 * Di - Du; zsel = 3 * WS; }
                    const float4* src = r4 + (row0 << g.logcpr()) + (c >> 2);
                    const int stride = F << g.logcpr();
                    float4 s = *reinterpret_cast<float4*>(acc + e);
#pragma unroll
                    for (int i = 0; i < K; ++i) {
                        const float4 v = src[i * stride];
                        const float z = wts[zsel + i];
                        s.x = fmaf(z, v.x, s.x); s.y = fmaf(z, v.y, s.y); s.z = fmaf(z, v.z, s.z); s.w = fmaf(z, v.w, s.w);
                    }
                    *reinterpret_cast<float4*>(acc + e) = s;
                }
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

template <class G>
static void coatt_bwd_launch(cudaStream_t st, const Dims& dm, const CoattBwdArgs& a) {
    const int warps = 4;
    size_t smem;
    CoattSmem sp = coatt_plan(dm, warps, true, &smem);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaFuncSetAttribute(coatt_bwd_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = smem;
    }
    launch_chain(coatt_bwd_kernel<G>, dim3(a.n_partials), dim3(warps * 32), smem, st, dm, a, sp);
}

void launch_coatt_bwd(cudaStream_t st, const Dims& dm, const CoattBwdArgs& a) {
    ++g_launch_count;
    if (try_coatt_bwd_lean(st, dm, a, a.live)) return;
    SCORE_GEOM_DISPATCH(coatt_bwd_launch, st, dm, a);
}

// ------------------------------------------------------------------------------------------
// target rows backward: one warp per sample, grid-stride, per-CTA partials for dWt / dbias.
__global__ void target_bwd_kernel(Dims dm, TargetBwdArgs a) {
    pdl_enter();
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

int target_bwd_num_ctas() { return 2 * num_sms(); }   // one sample per warp at B = 1024

void launch_target_bwd(cudaStream_t st, const Dims& dm, const TargetBwdArgs& a) {
    const int warps = 4;
    size_t smem = (size_t)warps * (dm.Di + dm.Du + 2) * sizeof(float);
    launch_chain(target_bwd_kernel, dim3(a.n_partials), dim3(warps * 32), smem, st, dm, a);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------
// co-attention kernel gradient [Wt | W1 | W2] + bias, for both co-attentions: one CTA of 128 threads per output
// element, every thread adds its rows of the per-CTA partials in index order (independent loads), then a fixed-shape
// tree over the CTA (deterministic).
__global__ void __launch_bounds__(128) coatt_grad_reduce_kernel(Dims dm, const float* __restrict__ cp, int n_coatt,
                                                                const float* __restrict__ tp, int n_target,
                                                                float* g_w_item, float* g_b_item, float* g_w_user, float* g_b_user) {
    pdl_enter();
    __shared__ float red[128];
    const int nacc_c = 2 * dm.Di + 2 * dm.Du;
    const int nacc_t = dm.Di + 1 + dm.Du + 1;
    const int e = blockIdx.x;
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
    for (int i = threadIdx.x; i < n; i += 128) s += src[(int64_t)i * stride + col];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *dst = red[0];
}

void launch_coatt_grad_reduce(cudaStream_t st, const Dims& dm, const float* coatt_partials, int n_coatt,
                              const float* target_partials, int n_target,
                              float* g_w_item, float* g_b_item, float* g_w_user, float* g_b_user) {
    int total = 3 * dm.Di + 1 + 3 * dm.Du + 1;
    launch_chain(coatt_grad_reduce_kernel, dim3(total), dim3(128), 0, st, dm, coatt_partials, n_coatt, target_partials, n_target, g_w_item,
                                                    g_b_item, g_w_user, g_b_user);
    ++g_launch_count;
}

}  // namespace score
