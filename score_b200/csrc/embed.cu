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
__global__ void build_keys_kernel(Dims dm, const int32_t* __restrict__ ids, const int32_t* __restrict__ length,
                                  int32_t* __restrict__ keys, int32_t* __restrict__ err_flag) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= dm.N) return;
    int32_t id = ids[p];
    bool live = true;
    if (p < dm.off_tu) {
        int64_t local; int f;
        if (p < dm.off_u2) { local = p - dm.off_u1; f = dm.fi; }
        else if (p < dm.off_i1) { local = p - dm.off_u2; f = dm.fu; }
        else if (p < dm.off_i2) { local = p - dm.off_i1; f = dm.fu; }
        else { local = p - dm.off_i2; f = dm.fi; }
        int64_t slice = local / ((int64_t)dm.K * f);
        int b = (int)(slice / dm.T), t = (int)(slice % dm.T);
        live = t < length[b];
    }
    if (id < 0 || (int64_t)id >= dm.V) { atomicExch(err_flag, 1); id = 0; }
    keys[p] = live ? id : 0;
}

void launch_build_keys(cudaStream_t st, const Dims& dm, const int32_t* ids, const int32_t* length,
                       int32_t* keys, int32_t* err_flag) {
    int threads = 256;
    int64_t blocks = (dm.N + threads - 1) / threads;
    build_keys_kernel<<<(unsigned)blocks, threads, 0, st>>>(dm, ids, length, keys, err_flag);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------
// stage `nrows` table rows (ids at key_ptr[0..nrows)) into smem dst[nrows][d]; whole warp cooperates,
// consecutive lanes fetch consecutive 16-byte chunks (a d=16 row is 4 lanes, 64 B contiguous).
__device__ __forceinline__ void stage_rows(float* dst, const float* __restrict__ emb, const int32_t* __restrict__ key_ptr,
                                           int nrows, int d, int lane) {
    const int cpr = d >> 2;   // 16-byte chunks per row
    const int total = nrows * cpr;
    for (int q = lane; q < total; q += 32) {
        int r = q / cpr, c = q - r * cpr;
        int32_t id = __ldg(key_ptr + r);
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
// shared-memory plan of the co-attention kernels (floats)
struct CoattSmem {
    int w_off;        // [3*Di + 3*Du] co-attention kernels (CTA-shared)
    int warp_off;     // per-warp region start
    int warp_stride;  // per-warp floats
    int rows;         // per-warp row staging floats: K*(2*Di + 2*Du)
    // within a warp region: rows | scratch
};
__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }

__device__ __forceinline__ void stage_slice(float* rows, const Dims& dm, const float* emb, const int32_t* keys,
                                            int64_t slice, int lane) {
    // layout: u1 [K*Di] | i2 [K*Di] | u2 [K*Du] | i1 [K*Du]
    const int KDi = dm.K * dm.Di, KDu = dm.K * dm.Du;
    stage_rows(rows, emb, keys + dm.off_u1 + slice * dm.K * dm.fi, dm.K * dm.fi, dm.d, lane);
    stage_rows(rows + KDi, emb, keys + dm.off_i2 + slice * dm.K * dm.fi, dm.K * dm.fi, dm.d, lane);
    stage_rows(rows + 2 * KDi, emb, keys + dm.off_u2 + slice * dm.K * dm.fu, dm.K * dm.fu, dm.d, lane);
    stage_rows(rows + 2 * KDi + KDu, emb, keys + dm.off_i1 + slice * dm.K * dm.fu, dm.K * dm.fu, dm.d, lane);
    cp_async_commit();
}

// one co-attention for one slice, rows already in smem.  W = [Wt | W1 | W2] (3*D floats).
// out1 = sum_i softmax(r)_i seq1[i];  out2 = mean_j seq2[j];  info[0:K] = K*r_i, info[K:2K] = sum r.
__device__ __forceinline__ void coatt_slice_fwd(const float* s1, const float* s2, int D, int K, const float* W, float cb,
                                                float* out1a, float* out1b, float* out2a, float* out2b, float* info,
                                                float* save_r, float* save_w, float* wbuf, int lane) {
    const float* W1 = W + D; const float* W2 = W + 2 * D;
    float my_r = 0.f;
    for (int i = 0; i < K; ++i) {
        float p = 0.f;
        for (int c = lane; c < D; c += 32) p += W1[c] * s1[i * D + c] + W2[c] * s2[i * D + c];
        p = warp_sum(p);
        if (lane == i) my_r = fmaxf(p + cb, 0.f);
    }
    float r = (lane < K) ? my_r : -INFINITY;
    float mx = warp_max(r);
    float e = (lane < K) ? expf(r - mx) : 0.f;
    float den = warp_sum(e);
    float w = e / den;
    float rs = warp_sum((lane < K) ? my_r : 0.f);
    if (lane < K) {
        info[lane] = (float)K * my_r;
        info[K + lane] = rs;
        save_r[lane] = my_r;
        save_w[lane] = w;
        wbuf[lane] = w;
    }
    __syncwarp();
    const float invK = 1.0f / (float)K;
    for (int c = lane; c < D; c += 32) {
        float a1 = 0.f, a2 = 0.f;
        for (int i = 0; i < K; ++i) { a1 += wbuf[i] * s1[i * D + c]; a2 += s2[i * D + c]; }
        a2 *= invK;
        out1a[c] = a1; out1b[c] = a1;
        out2a[c] = a2; out2b[c] = a2;
    }
    __syncwarp();
}

__global__ void coatt_fwd_kernel(Dims dm, CoattArgs a, CoattSmem sp) {
    extern __shared__ __align__(16) float sm[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Wsm = sm + sp.w_off;
    for (int i = threadIdx.x; i < 3 * dm.Di; i += blockDim.x) Wsm[i] = a.w_item[i];
    for (int i = threadIdx.x; i < 3 * dm.Du; i += blockDim.x) Wsm[3 * dm.Di + i] = a.w_user[i];
    __syncthreads();
    float* rows = sm + sp.warp_off + warp * sp.warp_stride;
    float* wbuf = rows + sp.rows;
    const int KDi = dm.K * dm.Di, KDu = dm.K * dm.Du;
    const int64_t M = (int64_t)dm.B * dm.T;
    for (int64_t slice = (int64_t)blockIdx.x * warps + warp; slice < M; slice += (int64_t)gridDim.x * warps) {
        const int b = (int)(slice / dm.T), t = (int)(slice % dm.T);
        float* xu_g = a.xhg_u + slice * dm.ldx; float* xu_c = a.xhc_u + slice * dm.ldx;
        float* xi_g = a.xhg_i + slice * dm.ldx; float* xi_c = a.xhc_i + slice * dm.ldx;
        float* info = a.key + slice * a.ldkey + a.key_off;
        if (t >= a.length[b]) {   // dead slice: nothing downstream reads it, keep buffers finite
            for (int c = lane; c < dm.Ds; c += 32) { xu_g[c] = 0.f; xu_c[c] = 0.f; xi_g[c] = 0.f; xi_c[c] = 0.f; }
            for (int c = lane; c < 4 * dm.K; c += 32) info[c] = 0.f;
            continue;
        }
        stage_slice(rows, dm, a.emb, a.keys, slice, lane);
        cp_async_wait<0>();
        __syncwarp();
        // co-attention #1: (user_1hop, item_2hop, target_item)  score.py:196
        //   user_side = [user_1hop_seq (Di) | user_2hop_seq (Du)],  item_side = [item_1hop_seq (Du) | item_2hop_seq (Di)]
        coatt_slice_fwd(rows, rows + KDi, dm.Di, dm.K, Wsm, a.c_item[b],
                        xu_g, xu_c, xi_g + dm.Du, xi_c + dm.Du, info,
                        a.save_r + slice * 2 * dm.K, a.save_w + slice * 2 * dm.K, wbuf, lane);
        // co-attention #2: (user_2hop, item_1hop, target_user)  score.py:197
        coatt_slice_fwd(rows + 2 * KDi, rows + 2 * KDi + KDu, dm.Du, dm.K, Wsm + 3 * dm.Di, a.c_user[b],
                        xu_g + dm.Di, xu_c + dm.Di, xi_g, xi_c, info + 2 * dm.K,
                        a.save_r + slice * 2 * dm.K + dm.K, a.save_w + slice * 2 * dm.K + dm.K, wbuf, lane);
    }
}

static CoattSmem coatt_plan(const Dims& dm, int warps, int scratch_floats, size_t* bytes) {
    CoattSmem sp;
    sp.w_off = 0;
    sp.warp_off = round4(3 * dm.Di + 3 * dm.Du);
    sp.rows = dm.K * (2 * dm.Di + 2 * dm.Du);
    sp.warp_stride = round4(sp.rows + scratch_floats);
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
    CoattSmem sp = coatt_plan(dm, warps, 32, &smem);
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
// backward of one co-attention for one slice.
//   dr_i = K*dinfo[i] + sum_j dinfo[K+j] + w_i (dw_i - sum_k w_k dw_k),  dw_i = dout1 . seq1[i]
//   dz_i = dr_i [r_i > 0]
//   dseq1[i] = w_i dout1 + dz_i W1      dseq2[i] = dout2 / K + dz_i W2
//   dW1 += sum_i dz_i seq1[i]           dW2 += sum_i dz_i seq2[i]       sdz = sum_i dz_i
__device__ __forceinline__ float coatt_slice_bwd(const float* s1, const float* s2, int D, int K, const float* W,
                                                 const float* dout1, const float* dout2, const float* dinfo,
                                                 const float* save_r, const float* save_w,
                                                 float* g1, float* g2,      // grad rows of seq1 / seq2: K*D floats each
                                                 float* accW1, float* accW2, // per-warp accumulators [D]
                                                 float* dzbuf, float* wbuf, float* d1buf, float* d2buf, int lane) {
    const float* W1 = W + D; const float* W2 = W + 2 * D;
    for (int c = lane; c < D; c += 32) { d1buf[c] = dout1[c]; d2buf[c] = dout2[c]; }
    __syncwarp();
    float my_dw = 0.f;
    for (int i = 0; i < K; ++i) {
        float p = 0.f;
        for (int c = lane; c < D; c += 32) p += d1buf[c] * s1[i * D + c];
        p = warp_sum(p);
        if (lane == i) my_dw = p;
    }
    float w = (lane < K) ? save_w[lane] : 0.f;
    float r = (lane < K) ? save_r[lane] : 0.f;
    float dot = warp_sum(w * my_dw);
    float tail = warp_sum((lane < K) ? dinfo[K + lane] : 0.f);
    float dr = (lane < K) ? ((float)K * dinfo[lane] + tail + w * (my_dw - dot)) : 0.f;
    float dz = (r > 0.f) ? dr : 0.f;
    float sdz = warp_sum(dz);
    if (lane < K) { dzbuf[lane] = dz; wbuf[lane] = w; }
    __syncwarp();
    const float invK = 1.0f / (float)K;
    const int total = K * D;
    for (int e = lane; e < total; e += 32) {
        int i = e / D, c = e - i * D;
        g1[e] = wbuf[i] * d1buf[c] + dzbuf[i] * W1[c];
        g2[e] = d2buf[c] * invK + dzbuf[i] * W2[c];
    }
    for (int c = lane; c < D; c += 32) {
        float a1 = 0.f, a2 = 0.f;
        for (int i = 0; i < K; ++i) { a1 += dzbuf[i] * s1[i * D + c]; a2 += dzbuf[i] * s2[i * D + c]; }
        accW1[c] += a1; accW2[c] += a2;
    }
    __syncwarp();
    return sdz;
}

__global__ void coatt_bwd_kernel(Dims dm, CoattBwdArgs a, CoattSmem sp) {
    extern __shared__ __align__(16) float sm[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Wsm = sm + sp.w_off;
    for (int i = threadIdx.x; i < 3 * dm.Di; i += blockDim.x) Wsm[i] = a.w_item[i];
    for (int i = threadIdx.x; i < 3 * dm.Du; i += blockDim.x) Wsm[3 * dm.Di + i] = a.w_user[i];
    float* rows = sm + sp.warp_off + warp * sp.warp_stride;
    const int Dmax = dm.Di > dm.Du ? dm.Di : dm.Du;
    const int nacc = 2 * dm.Di + 2 * dm.Du;
    float* acc = rows + sp.rows;            // [2*Di + 2*Du] dW1_item | dW2_item | dW1_user | dW2_user
    float* dzbuf = acc + nacc;              // [32]
    float* wbuf = dzbuf + 32;               // [32]
    float* d1buf = wbuf + 32;               // [Dmax]
    float* d2buf = d1buf + Dmax;            // [Dmax]
    for (int c = lane; c < nacc; c += 32) acc[c] = 0.f;
    __syncthreads();
    const int KDi = dm.K * dm.Di, KDu = dm.K * dm.Du;
    const int64_t M = (int64_t)dm.B * dm.T;
    for (int64_t slice = (int64_t)blockIdx.x * warps + warp; slice < M; slice += (int64_t)gridDim.x * warps) {
        const int b = (int)(slice / dm.T), t = (int)(slice % dm.T);
        if (t >= a.length[b]) {
            if (lane == 0) { a.sdz[slice * 2] = 0.f; a.sdz[slice * 2 + 1] = 0.f; }
            continue;   // positions of dead slices carry key 0: their gradient rows are never read
        }
        stage_slice(rows, dm, a.emb, a.keys, slice, lane);
        cp_async_wait<0>();
        __syncwarp();
        const float* dxu = a.dxu + slice * dm.Ds; const float* dxi = a.dxi + slice * dm.Ds;
        const float* dinfo = a.dkey + slice * a.ldkey + a.key_off;
        float* gr = a.grad_rows;
        float s_item = coatt_slice_bwd(rows, rows + KDi, dm.Di, dm.K, Wsm,
                                       dxu, dxi + dm.Du, dinfo,
                                       a.save_r + slice * 2 * dm.K, a.save_w + slice * 2 * dm.K,
                                       gr + (dm.off_u1 + slice * dm.K * dm.fi) * dm.d,
                                       gr + (dm.off_i2 + slice * dm.K * dm.fi) * dm.d,
                                       acc, acc + dm.Di, dzbuf, wbuf, d1buf, d2buf, lane);
        float s_user = coatt_slice_bwd(rows + 2 * KDi, rows + 2 * KDi + KDu, dm.Du, dm.K, Wsm + 3 * dm.Di,
                                       dxu + dm.Di, dxi, dinfo + 2 * dm.K,
                                       a.save_r + slice * 2 * dm.K + dm.K, a.save_w + slice * 2 * dm.K + dm.K,
                                       gr + (dm.off_u2 + slice * dm.K * dm.fu) * dm.d,
                                       gr + (dm.off_i1 + slice * dm.K * dm.fu) * dm.d,
                                       acc + 2 * dm.Di, acc + 2 * dm.Di + dm.Du, dzbuf, wbuf, d1buf, d2buf, lane);
        if (lane == 0) { a.sdz[slice * 2] = s_item; a.sdz[slice * 2 + 1] = s_user; }
    }
    __syncthreads();
    // fixed-order sum over the CTA's warps -> one partial row per CTA
    for (int c = threadIdx.x; c < nacc; c += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < warps; ++w) s += sm[sp.warp_off + w * sp.warp_stride + sp.rows + c];
        a.partials[(int64_t)blockIdx.x * nacc + c] = s;
    }
}

int coatt_bwd_num_ctas() { return num_sms() * 2; }

void launch_coatt_bwd(cudaStream_t st, const Dims& dm, const CoattBwdArgs& a) {
    const int warps = 4;
    const int Dmax = dm.Di > dm.Du ? dm.Di : dm.Du;
    size_t smem;
    CoattSmem sp = coatt_plan(dm, warps, 2 * dm.Di + 2 * dm.Du + 64 + 2 * Dmax, &smem);
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
// co-attention kernel gradient [Wt | W1 | W2] + bias, for both co-attentions, fixed-order sums.
__global__ void coatt_grad_reduce_kernel(Dims dm, const float* __restrict__ cp, int n_coatt,
                                         const float* __restrict__ tp, int n_target,
                                         float* g_w_item, float* g_b_item, float* g_w_user, float* g_b_user) {
    const int nacc_c = 2 * dm.Di + 2 * dm.Du;
    const int nacc_t = dm.Di + 1 + dm.Du + 1;
    const int total = 3 * dm.Di + 1 + 3 * dm.Du + 1;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
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
        for (int i = 0; i < n; ++i) s += src[(int64_t)i * stride + col];
        *dst = s;
    }
}

void launch_coatt_grad_reduce(cudaStream_t st, const Dims& dm, const float* coatt_partials, int n_coatt,
                              const float* target_partials, int n_target,
                              float* g_w_item, float* g_b_item, float* g_w_user, float* g_b_user) {
    int total = 3 * dm.Di + 1 + 3 * dm.Du + 1;
    coatt_grad_reduce_kernel<<<(total + 127) / 128, 128, 0, st>>>(dm, coatt_partials, n_coatt, target_partials,
                                                                   n_target, g_w_item, g_b_item, g_w_user, g_b_user);
    ++g_launch_count;
}

}  // namespace score
