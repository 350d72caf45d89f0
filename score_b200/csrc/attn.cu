// attention() over the T time slices (score.py:169-186) + attentive pooling (score.py:214-216), forward and
// backward, as fused row-tile chains on the tile_layer primitive (tile.cuh).
//
// The reference materialises inp = [q | key | q-key | q*key] ([B,T,4*Dk]) and multiplies it by dense_3/kernel
// ([4*Dk, 80], row blocks Wa | Wb | Wc | Wd).  Algebraically
//     inp W1 = q (Wa + Wc)  +  key (Wb - Wc)  +  (q*key) Wd
// and the first term depends on the sample only, not on t.  So:
//   att_q_kernel    (B rows)    q = q0 Wq + bq ;  U = q (Wa+Wc) + b1
//   att_fwd2_kernel (G samples = G*T rows per CTA)  f1 = relu(U[b] + [key | q*key] [Wb-Wc ; Wd]) -> f2 -> raw score ->
//                   masked softmax over T -> pooled user/item state.  Nothing of width 4*Dk exists.
//   att_bwd2_kernel pooling + softmax backward -> df2 -> df1 -> dV = df1 (Wb-Wc)^T, dD = df1 Wd^T ->
//                   dkey = dV + dD*q (+ pooling gradient);  per sample  sdf1 = sum_t df1,  dqD = sum_t dD*key
//   att_qb_kernel   (B rows)    dq = sdf1 (Wa+Wc)^T + dqD ;  dq0 = dq Wq^T
// The derived matrices (Wb-Wc etc. and the transposes the backward chains stream) are rebuilt once per step by
// prep_weights_kernel.  Weight gradients: dWa = q^T sdf1, dWb = key^T df1, dWd = (q*key)^T df1 as split GEMMs
// (gemm.cu); dWc = dWa - dWb is formed in reduce_partials.
#include "kernels.h"
#include "tile.cuh"

namespace score {

namespace {

constexpr int A1 = 80, A2 = 40;   // widths of the attention MLP (score.py:175-176)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float pick(const float4& v, int j) { return j == 0 ? v.x : j == 1 ? v.y : j == 2 ? v.z : v.w; }

}  // namespace

// ------------------------------------------------------------------------------------------ derived weights
__global__ void prep_weights_kernel(PrepOps ops, const float* __restrict__ P, float* __restrict__ D) {
    const PrepOp& o = ops.op[blockIdx.y];
    const int n = o.rows * o.cols;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / o.cols, c = i - r * o.cols;
        float v = P[o.a_off + (int64_t)r * o.a_rs + (int64_t)c * o.a_cs];
        if (o.b_off >= 0) {
            const float w = P[o.b_off + (int64_t)r * o.a_rs + (int64_t)c * o.a_cs];
            v = o.sign > 0 ? v + w : v - w;
        }
        D[o.dst_off + (int64_t)r * o.ld_dst + c] = v;
    }
}
void launch_prep_weights(cudaStream_t st, const PrepOps& ops, const float* P, float* D) {
    if (ops.n <= 0) return;
    prep_weights_kernel<<<dim3(16, ops.n), 256, 0, st>>>(ops, P, D);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------ per-sample query side
__global__ void __launch_bounds__(TL_CT) att_q_kernel(AttQArgs a) {
    constexpr int RT = 16;
    using G = TileGeom<RT>;
    extern __shared__ __align__(16) float sm[];
    const int Ds = a.Ds, Dk = a.Dk, tid = threadIdx.x;
    float* X0 = sm;                  // [Ds][RT]
    float* Xq = X0 + Ds * RT;        // [Dk][RT]
    float* wbuf = Xq + Dk * RT;
    const int row0 = blockIdx.x * RT;
    int rg, cg;
    tile_coords<RT>(tid, rg, cg);
    tile_for_each_chunk<RT>(Ds, tid, [&](int r, int c4) {
        const int b = row0 + r;
        const float4 v = b < a.B ? ld4(a.q0 + (int64_t)b * Ds + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        tile_put4(X0, RT, 4 * c4, r, v);
    });
    __syncthreads();
    float acc[4][4];
    for (int cb = 0; cb < Dk; cb += G::NP) {
        const int nb = min(G::NP, Dk - cb);
        tile_zero(acc);
        tile_layer<RT>(X0, Ds, a.wq + cb, Dk, nb, acc, wbuf, tid);
        if (4 * cg < nb) {
            const int c = cb + 4 * cg;
            const float4 bias = ld4(a.bq + c);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] += bias.x; acc[i][1] += bias.y; acc[i][2] += bias.z; acc[i][3] += bias.w;
                const int b = row0 + 4 * rg + i;
                if (b < a.B) st4(a.q + (int64_t)b * Dk + c, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
            }
            tile_store_smem<RT>(Xq + cb * RT, acc, rg, cg);
        }
    }
    __syncthreads();
    tile_zero(acc);
    tile_layer<RT>(Xq, Dk, a.Wac, A1, A1, acc, wbuf, tid);
    if (cg < A1 / 4) {
        const float4 bias = ld4(a.b1 + 4 * cg);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = row0 + 4 * rg + i;
            if (b < a.B)
                st4(a.U + (int64_t)b * A1 + 4 * cg,
                    make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w));
        }
    }
}

void launch_att_q(cudaStream_t st, const AttQArgs& a) {
    constexpr int RT = 16;
    const size_t smem = ((size_t)(a.Ds + a.Dk) * RT + TileGeom<RT>::WBUF) * sizeof(float);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaFuncSetAttribute(att_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = smem;
    }
    att_q_kernel<<<(a.B + RT - 1) / RT, TL_CT, smem, st>>>(a);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------ attention forward
// shared-memory plan of att_fwd2 / att_bwd2 (floats); rows_pad = G*T rounded up to 32
struct AttPlan { int rows_pad; size_t bytes; };
static AttPlan att_fwd_plan(int Dk, int T, int G) {
    AttPlan p;
    p.rows_pad = (G * T + 31) & ~31;
    size_t fl = (size_t)2 * Dk * 32 + A1 * 32 + A2 * 32 + p.rows_pad + 64 + TileGeom<32>::WBUF;
    p.bytes = fl * sizeof(float);
    return p;
}

__global__ void __launch_bounds__(TL_CT) att_fwd2_kernel(AttFwd2Args a, int rows_pad) {
    pdl_enter();
    constexpr int RT = 32;
    extern __shared__ __align__(16) float sm[];
    const int Dk = a.Dk, T = a.T, G = a.G, H = a.H, tid = threadIdx.x;
    const int rows = G * T;
    float* Xs = sm;                          // [2*Dk][RT]   key | q*key
    float* X1 = Xs + 2 * Dk * RT;            // [80][RT]
    float* X2 = X1 + A1 * RT;                // [40][RT]
    float* sc = X2 + A2 * RT;                // [rows_pad]   raw scores, then softmax weights
    int* rowm = reinterpret_cast<int*>(sc + rows_pad);   // [RT] global row of tile row r, -1: none
    int* rowg = rowm + RT;                   // [RT] local sample of tile row r
    float* wbuf = reinterpret_cast<float*>(rowg + RT);
    const int b0 = blockIdx.x * G;
    int rg, cg;
    tile_coords<RT>(tid, rg, cg);
    float acc[4][4];
    for (int r0 = 0; r0 < rows; r0 += RT) {
        if (tid < RT) {
            const int lr = r0 + tid;
            int m = -1, g = 0;
            if (lr < rows) {
                g = lr / T;
                if (b0 + g < a.B) m = (b0 + g) * T + (lr - g * T);
            }
            rowm[tid] = m; rowg[tid] = g;
        }
        __syncthreads();
        tile_for_each_chunk<RT>(Dk, tid, [&](int r, int c4) {
            const int m = rowm[r];
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), pv = kv;
            if (m >= 0) {
                kv = ld4(a.key + (int64_t)m * Dk + 4 * c4);
                const float4 qv = ld4(a.q + (int64_t)(b0 + rowg[r]) * Dk + 4 * c4);
                pv = make_float4(kv.x * qv.x, kv.y * qv.y, kv.z * qv.z, kv.w * qv.w);
                st4(a.qk + (int64_t)m * Dk + 4 * c4, pv);
            }
            tile_put4(Xs, RT, 4 * c4, r, kv);
            tile_put4(Xs + Dk * RT, RT, 4 * c4, r, pv);
        });
        __syncthreads();
        // layer 1: f1 = relu(U[b] + [key | q*key] [Wb-Wc ; Wd])
        tile_zero(acc);
        tile_layer<RT>(Xs, 2 * Dk, a.W1e, A1, A1, acc, wbuf, tid);
        if (cg < A1 / 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = 4 * rg + i, m = rowm[r];
                if (m >= 0) {
                    const float4 u = ld4(a.U + (int64_t)(b0 + rowg[r]) * A1 + 4 * cg);
                    acc[i][0] = fmaxf(acc[i][0] + u.x, 0.f); acc[i][1] = fmaxf(acc[i][1] + u.y, 0.f);
                    acc[i][2] = fmaxf(acc[i][2] + u.z, 0.f); acc[i][3] = fmaxf(acc[i][3] + u.w, 0.f);
                    st4(a.f1 + (int64_t)m * A1 + 4 * cg, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
                } else {
                    acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
                }
            }
            tile_store_smem<RT>(X1, acc, rg, cg);
        }
        __syncthreads();
        // layer 2: f2 = relu(f1 W2 + b2)
        tile_zero(acc);
        tile_layer<RT>(X1, A1, a.w2, A2, A2, acc, wbuf, tid);
        if (cg < A2 / 4) {
            const float4 bias = ld4(a.b2 + 4 * cg);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = 4 * rg + i, m = rowm[r];
                if (m >= 0) {
                    acc[i][0] = fmaxf(acc[i][0] + bias.x, 0.f); acc[i][1] = fmaxf(acc[i][1] + bias.y, 0.f);
                    acc[i][2] = fmaxf(acc[i][2] + bias.z, 0.f); acc[i][3] = fmaxf(acc[i][3] + bias.w, 0.f);
                    st4(a.f2 + (int64_t)m * A2 + 4 * cg, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
                } else {
                    acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
                }
            }
            tile_store_smem<RT>(X2, acc, rg, cg);
        }
        __syncthreads();
        // layer 3: raw score
        if (tid < RT && rowm[tid] >= 0) {
            float s = 0.f;
#pragma unroll 8
            for (int n = 0; n < A2; ++n) s = fmaf(X2[n * RT + tid], a.w3[n], s);
            sc[r0 + tid] = s + a.b3[0];
        }
    }
    __syncthreads();
    // masked softmax over T (padding -2**32+1, score.py:179-181) and attentive pooling: one warp per sample
    const int warp = tid >> 5, lane = tid & 31;
    const float pad = -4294967296.0f;
    for (int g = warp; g < G; g += TL_CT / 32) {
        const int b = b0 + g;
        if (b >= a.B) continue;
        const int len = a.length[b];
        float* s = sc + g * T;
        float mx = -INFINITY;
        for (int t = lane; t < T; t += 32) {
            const float v = (t < len) ? s[t] : pad;
            s[t] = v;
            mx = fmaxf(mx, v);
        }
        mx = warp_max(mx);
        float den = 0.f;
        for (int t = lane; t < T; t += 32) den += expf(s[t] - mx);
        den = warp_sum(den);
        for (int t = lane; t < T; t += 32) {
            const float w = expf(s[t] - mx) / den;
            s[t] = w;
            a.score[(int64_t)b * T + t] = w;
        }
        __syncwarp();
        const int mt = a.model_type;
        for (int c = lane; c < 2 * H; c += 32) {
            float p = 0.f;
            for (int t = 0; t < T; ++t) p += a.key[((int64_t)b * T + t) * Dk + c] * s[t];
            int dst = c;
            if (mt == 3) { if (c >= H) continue; }
            else if (mt == 4) { if (c < H) continue; dst = c - H; }
            a.fc_in[(int64_t)b * a.ldfc + dst] = p;
        }
    }
}

static int att_group(int T) { return T >= 32 ? 1 : 32 / T; }

void launch_att_fwd2(cudaStream_t st, AttFwd2Args a) {
    a.G = att_group(a.T);
    const AttPlan p = att_fwd_plan(a.Dk, a.T, a.G);
    static size_t attr_set = 0;
    if (p.bytes > 48 * 1024 && p.bytes > attr_set) {
        cudaFuncSetAttribute(att_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.bytes);
        attr_set = p.bytes;
    }
    launch_chain(att_fwd2_kernel, dim3((a.B + a.G - 1) / a.G), dim3(TL_CT), p.bytes, st, a, p.rows_pad);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------ attention backward
struct AttBwdPlan { int rows_pad, dfin, sdf, dqs; size_t bytes; };
static AttBwdPlan att_bwd_plan(int Dk, int T, int G, int H) {
    AttBwdPlan p;
    p.rows_pad = (G * T + 31) & ~31;
    p.dfin = (G * 2 * H + 3) & ~3;
    p.sdf = G * A1;
    p.dqs = (G * Dk + 3) & ~3;
    size_t fl = (size_t)A2 * 32 + A1 * 32 + 128 * 32 + 2 * p.rows_pad + p.dfin + p.sdf + p.dqs + 64 + TileGeom<32>::WBUF;
    p.bytes = fl * sizeof(float);
    return p;
}

__global__ void __launch_bounds__(TL_CT) att_bwd2_kernel(AttBwd2Args a, AttBwdPlan pl) {
    pdl_enter();
    constexpr int RT = 32;
    using GE = TileGeom<RT>;
    extern __shared__ __align__(16) float sm[];
    const int Dk = a.Dk, T = a.T, G = a.G, H = a.H, H2 = 2 * a.H, tid = threadIdx.x;
    const int rows = G * T;
    float* D2 = sm;                          // [40][RT]   d f2
    float* D1 = D2 + A2 * RT;                // [80][RT]   d f1
    float* DQ = D1 + A1 * RT;                // [128][RT]  dD * key of the current column block
    float* dsr = DQ + 128 * RT;              // [rows_pad] d raw score
    float* scr = dsr + pl.rows_pad;          // [rows_pad] softmax weight
    float* dfin = scr + pl.rows_pad;         // [G][2H]    d pooled state (user | item)
    float* sdf = dfin + pl.dfin;             // [G][80]    sum_t d f1
    float* dqs = sdf + pl.sdf;               // [G][Dk]    sum_t dD * key
    int* rowm = reinterpret_cast<int*>(dqs + pl.dqs);
    int* rowg = rowm + RT;
    float* wbuf = reinterpret_cast<float*>(rowg + RT);
    const int b0 = blockIdx.x * G;
    const int warp = tid >> 5, lane = tid & 31;
    // phase 0: pooling + softmax backward, one warp per sample
    for (int g = warp; g < G; g += TL_CT / 32) {
        const int b = b0 + g;
        if (b >= a.B) {
            for (int t = lane; t < T; t += 32) { dsr[g * T + t] = 0.f; scr[g * T + t] = 0.f; }
            for (int c = lane; c < H2; c += 32) dfin[g * H2 + c] = 0.f;
            continue;
        }
        const int len = a.length[b], mt = a.model_type;
        const float* dfc = a.dfc_in + (int64_t)b * a.ldfc;
        for (int c = lane; c < H2; c += 32) {
            float d;
            if (mt == 3) d = (c < H) ? dfc[c] : 0.f;
            else if (mt == 4) d = (c >= H) ? dfc[c - H] : 0.f;
            else d = dfc[c];
            dfin[g * H2 + c] = d;
        }
        __syncwarp();
        for (int t = 0; t < T; ++t) {
            const float* kr = a.key + ((int64_t)b * T + t) * Dk;
            float p = 0.f;
            for (int c = lane; c < H2; c += 32) p += dfin[g * H2 + c] * kr[c];
            p = warp_sum(p);
            if (lane == 0) dsr[g * T + t] = p;
        }
        __syncwarp();
        float dot = 0.f;
        for (int t = lane; t < T; t += 32) dot += a.score[(int64_t)b * T + t] * dsr[g * T + t];
        dot = warp_sum(dot);
        for (int t = lane; t < T; t += 32) {
            const float s = a.score[(int64_t)b * T + t];
            const float v = (t < len) ? s * (dsr[g * T + t] - dot) : 0.f;   // tf.where blocks the gradient of padded slots
            dsr[g * T + t] = v;
            scr[g * T + t] = s;
            a.ds[(int64_t)b * T + t] = v;
        }
    }
    for (int i = tid; i < pl.sdf; i += TL_CT) sdf[i] = 0.f;
    for (int i = tid; i < pl.dqs; i += TL_CT) dqs[i] = 0.f;
    int rg, cg;
    tile_coords<RT>(tid, rg, cg);
    float acc[4][4], acd[4][4];
    for (int r0 = 0; r0 < rows; r0 += RT) {
        if (tid < RT) {
            const int lr = r0 + tid;
            int m = -1, g = 0;
            if (lr < rows) {
                g = lr / T;
                if (b0 + g < a.B) m = (b0 + g) * T + (lr - g * T);
            }
            rowm[tid] = m; rowg[tid] = g;
        }
        __syncthreads();
        // d f2 = ds * w3 (*) [f2 > 0]
        tile_for_each_chunk<RT>(A2, tid, [&](int r, int c4) {
            const int m = rowm[r];
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m >= 0) {
                const float4 f = ld4(a.f2 + (int64_t)m * A2 + 4 * c4);
                const float4 w = ld4(a.w3 + 4 * c4);
                const float d = dsr[r0 + r];
                v = make_float4(f.x > 0.f ? d * w.x : 0.f, f.y > 0.f ? d * w.y : 0.f, f.z > 0.f ? d * w.z : 0.f,
                                f.w > 0.f ? d * w.w : 0.f);
                st4(a.df2 + (int64_t)m * A2 + 4 * c4, v);
            }
            tile_put4(D2, RT, 4 * c4, r, v);
        });
        __syncthreads();
        // d f1 = (d f2 W2^T) (*) [f1 > 0]
        tile_zero(acc);
        tile_layer<RT>(D2, A2, a.W2T, A1, A1, acc, wbuf, tid);
        if (cg < A1 / 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = 4 * rg + i, m = rowm[r];
                if (m >= 0) {
                    const float4 f = ld4(a.f1 + (int64_t)m * A1 + 4 * cg);
                    acc[i][0] = f.x > 0.f ? acc[i][0] : 0.f; acc[i][1] = f.y > 0.f ? acc[i][1] : 0.f;
                    acc[i][2] = f.z > 0.f ? acc[i][2] : 0.f; acc[i][3] = f.w > 0.f ? acc[i][3] : 0.f;
                    st4(a.df1 + (int64_t)m * A1 + 4 * cg, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
                } else {
                    acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
                }
            }
            tile_store_smem<RT>(D1, acc, rg, cg);
        }
        __syncthreads();
        // per-sample sum of d f1 over the sample's rows of this tile (fixed order)
        for (int idx = tid; idx < G * A1; idx += TL_CT) {
            const int g = idx % G, n = idx / G;
            const int lo = max(g * T, r0) - r0, hi = min(g * T + T, r0 + RT) - r0;
            float s = 0.f;
            for (int lr = lo; lr < hi; ++lr) s += D1[n * RT + lr];
            if (hi > lo) sdf[g * A1 + n] += s;
        }
        // d key and the key side of d q, 128 key columns at a time
        for (int cb = 0; cb < Dk; cb += GE::NP) {
            const int nb = min(GE::NP, Dk - cb);
            tile_zero(acc);
            tile_layer<RT>(D1, A1, a.W1eT + cb, 2 * Dk, nb, acc, wbuf, tid);        // dV = df1 (Wb-Wc)^T
            tile_zero(acd);
            tile_layer<RT>(D1, A1, a.W1eT + Dk + cb, 2 * Dk, nb, acd, wbuf, tid);   // dD = df1 Wd^T
            if (4 * cg < nb) {
                const int c = cb + 4 * cg;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = 4 * rg + i, m = rowm[r];
                    if (m >= 0) {
                        const int g = rowg[r];
                        const float4 qv = ld4(a.q + (int64_t)(b0 + g) * Dk + c);
                        const float4 kv = ld4(a.key + (int64_t)m * Dk + c);
                        const float sw = scr[r0 + r];
                        float dk[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            dk[j] = acc[i][j] + acd[i][j] * pick(qv, j);
                            if (c + j < H2) dk[j] += sw * dfin[g * H2 + c + j];   // pooling gradient of rep_t
                            acd[i][j] *= pick(kv, j);
                        }
                        st4(a.dkey + (int64_t)m * Dk + c, make_float4(dk[0], dk[1], dk[2], dk[3]));
                    } else {
                        acd[i][0] = acd[i][1] = acd[i][2] = acd[i][3] = 0.f;
                    }
                }
                tile_store_smem<RT>(DQ, acd, rg, cg);
            }
            __syncthreads();
            for (int idx = tid; idx < G * nb; idx += TL_CT) {
                const int g = idx % G, c = idx / G;
                const int lo = max(g * T, r0) - r0, hi = min(g * T + T, r0 + RT) - r0;
                float s = 0.f;
                for (int lr = lo; lr < hi; ++lr) s += DQ[c * RT + lr];
                if (hi > lo) dqs[g * Dk + cb + c] += s;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    for (int idx = tid; idx < G * A1; idx += TL_CT) {
        const int g = idx / A1, n = idx - g * A1;
        if (b0 + g < a.B) a.sdf1[(int64_t)(b0 + g) * A1 + n] = sdf[idx];
    }
    for (int idx = tid; idx < G * Dk; idx += TL_CT) {
        const int g = idx / Dk, c = idx - g * Dk;
        if (b0 + g < a.B) a.dqD[(int64_t)(b0 + g) * Dk + c] = dqs[idx];
    }
}

void launch_att_bwd2(cudaStream_t st, AttBwd2Args a) {
    a.G = att_group(a.T);
    const AttBwdPlan p = att_bwd_plan(a.Dk, a.T, a.G, a.H);
    static size_t attr_set = 0;
    if (p.bytes > 48 * 1024 && p.bytes > attr_set) {
        cudaFuncSetAttribute(att_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.bytes);
        attr_set = p.bytes;
    }
    launch_chain(att_bwd2_kernel, dim3((a.B + a.G - 1) / a.G), dim3(TL_CT), p.bytes, st, a, p);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------ per-sample query side, backward
__global__ void __launch_bounds__(TL_CT) att_qb_kernel(AttQbArgs a) {
    pdl_enter();
    constexpr int RT = 16;
    using G = TileGeom<RT>;
    extern __shared__ __align__(16) float sm[];
    const int Ds = a.Ds, Dk = a.Dk, tid = threadIdx.x;
    float* Xs = sm;                   // [80][RT]  sum_t d f1
    float* Xd = Xs + A1 * RT;         // [Dk][RT]  d q
    float* wbuf = Xd + Dk * RT;
    const int row0 = blockIdx.x * RT;
    int rg, cg;
    tile_coords<RT>(tid, rg, cg);
    tile_for_each_chunk<RT>(A1, tid, [&](int r, int c4) {
        const int b = row0 + r;
        const float4 v = b < a.B ? ld4(a.sdf1 + (int64_t)b * A1 + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        tile_put4(Xs, RT, 4 * c4, r, v);
    });
    __syncthreads();
    float acc[4][4];
    for (int cb = 0; cb < Dk; cb += G::NP) {
        const int nb = min(G::NP, Dk - cb);
        tile_zero(acc);
        tile_layer<RT>(Xs, A1, a.WacT + cb, Dk, nb, acc, wbuf, tid);
        if (4 * cg < nb) {
            const int c = cb + 4 * cg;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = row0 + 4 * rg + i;
                if (b < a.B) {
                    const float4 d = ld4(a.dqD + (int64_t)b * Dk + c);
                    acc[i][0] += d.x; acc[i][1] += d.y; acc[i][2] += d.z; acc[i][3] += d.w;
                    st4(a.dq + (int64_t)b * Dk + c, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
                } else {
                    acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
                }
            }
            tile_store_smem<RT>(Xd + cb * RT, acc, rg, cg);
        }
    }
    __syncthreads();
    for (int cb = 0; cb < Ds; cb += G::NP) {
        const int nb = min(G::NP, Ds - cb);
        tile_zero(acc);
        tile_layer<RT>(Xd, Dk, a.WqT + cb, Ds, nb, acc, wbuf, tid);
        if (4 * cg < nb) {
            const int c = cb + 4 * cg;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = row0 + 4 * rg + i;
                if (b < a.B) st4(a.dq0 + (int64_t)b * Ds + c, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
            }
        }
    }
}

void launch_att_qb(cudaStream_t st, const AttQbArgs& a) {
    constexpr int RT = 16;
    const size_t smem = ((size_t)(A1 + a.Dk) * RT + TileGeom<RT>::WBUF) * sizeof(float);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaFuncSetAttribute(att_qb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = smem;
    }
    launch_chain(att_qb_kernel, dim3((a.B + RT - 1) / RT), dim3(TL_CT), smem, st, a);
    ++g_launch_count;
}

}  // namespace score

// ------------------------------------------------------------------------------------------ generic row-tile product
// out[m, :N] (+)= X[m, :K] W[:K, :N] (+ bias) for up to 4 independent problems in one launch (blockIdx.y): the GRU input
// projections of both sides (score.py:205-208, the x part of [x, h] W hoisted out of the time loop) and their
// backward dx = dpx [Wg_x | Wc_x]^T.  32 rows per CTA on tile_layer.
namespace score {

__global__ void __launch_bounds__(TL_CT) rowgemm_kernel(RowGemmBatch batch) {
    pdl_enter();
    constexpr int RT = 32;
    using G = TileGeom<RT>;
    extern __shared__ __align__(16) float sm[];
    const RowGemmArgs& p = batch.p[blockIdx.y];
    const int tid = threadIdx.x, K = p.K;
    const int row0 = blockIdx.x * RT;
    if (row0 >= p.M) return;
    float* Xs = sm;               // [K][RT]
    float* wbuf = Xs + K * RT;
    int rg, cg;
    tile_coords<RT>(tid, rg, cg);
    tile_for_each_chunk<RT>(K, tid, [&](int r, int c4) {
        const int m = row0 + r;
        const float4 v = m < p.M ? *reinterpret_cast<const float4*>(p.X + (int64_t)m * p.ldx + 4 * c4)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
        tile_put4(Xs, RT, 4 * c4, r, v);
    });
    __syncthreads();
    float acc[4][4];
    for (int cb = 0; cb < p.N; cb += G::NP) {
        const int nb = min(G::NP, p.N - cb);
        tile_zero(acc);
        tile_layer<RT>(Xs, K, p.W + cb, p.ldw, nb, acc, wbuf, tid);
        if (4 * cg < nb) {
            const int c = cb + 4 * cg;
            float4 bias = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) bias = *reinterpret_cast<const float4*>(p.bias + c);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = row0 + 4 * rg + i;
                if (m < p.M)
                    *reinterpret_cast<float4*>(p.out + (int64_t)m * p.ldo + c) =
                        make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w);
            }
        }
    }
}

void launch_rowgemm(cudaStream_t st, const RowGemmArgs* list, int n) {
    if (n <= 0) return;
    RowGemmBatch b{};
    int maxM = 0, maxK = 0;
    for (int i = 0; i < n && i < 4; ++i) { b.p[i] = list[i]; maxM = max(maxM, list[i].M); maxK = max(maxK, list[i].K); }
    const size_t smem = ((size_t)maxK * 32 + TileGeom<32>::WBUF) * sizeof(float);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaFuncSetAttribute(rowgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = smem;
    }
    launch_chain(rowgemm_kernel, dim3((maxM + 31) / 32, n), dim3(TL_CT), smem, st, b);
    ++g_launch_count;
}

}  // namespace score
