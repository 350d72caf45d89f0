// Prediction head of the path as fused row-tile chains (build_fc_net + log_loss, score.py:68-81):
// BN (inference mode) -> 200 relu dropout -> 80 relu dropout -> 1 -> sigmoid -> loss, and its backward data chain.
// One CTA carries a tile of 16 samples through the whole chain on the tile_layer primitive (tile.cuh): activations
// stay in shared memory, weights stream through a cp.async ring, 4x4 register micro-tiles.  The backward chain
// streams the transposed kernels prep_weights_kernel (attn.cu) rebuilds every step; the weight gradients stay
// split GEMMs on the side stream.
#include "kernels.h"
#include "tile.cuh"

namespace score {

namespace {
constexpr int R = 16;          // samples per CTA
constexpr int F1 = 200, F2 = 80;   // widths of fc1 / fc2 (score.py:70-72)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
}  // namespace

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(TL_CT) fc_fwd_kernel(FcArgs a) {
    pdl_enter();
    extern __shared__ __align__(16) float sm[];
    const int F = a.F, tid = threadIdx.x;
    float* X0 = sm;                 // [F][R]   bn1 output
    float* X1 = X0 + F * R;         // [200][R] fc1 output (after dropout)
    float* X2 = X1 + F1 * R;        // [80][R]  fc2 output (after dropout)
    float* WB = X2 + F2 * R;        // weight staging
    const int row0 = blockIdx.x * R;
    const float keep = a.hp->keep_prob;
    const bool drop = (a.hp->train != 0) && keep < 1.f;
    const uint32_t s_lo = a.hp->seed_lo, s_hi = a.hp->seed_hi, step = (uint32_t)a.hp->step;
    const uint64_t sbase = (uint64_t)(uint32_t)a.hp->sample_base;   // dropout masks are keyed by the GLOBAL sample index
    int rg, cg;
    tile_coords<R>(tid, rg, cg);
    // batch norm in inference mode (score.py:69)
    tile_for_each_chunk<R>(F, tid, [&](int r, int c4) {
        const int gm = row0 + r, c = 4 * c4;
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gm < a.B) {
            const float4 x = ld4(a.fc_in + (int64_t)gm * F + c), ga = ld4(a.gamma + c), be = ld4(a.beta + c),
                         mu = ld4(a.mean + c), va = ld4(a.var + c);
            float inv;
            inv = ga.x / sqrtf(va.x + 1e-3f); z.x = x.x * inv + (be.x - mu.x * inv);
            inv = ga.y / sqrtf(va.y + 1e-3f); z.y = x.y * inv + (be.y - mu.y * inv);
            inv = ga.z / sqrtf(va.z + 1e-3f); z.z = x.z * inv + (be.z - mu.z * inv);
            inv = ga.w / sqrtf(va.w + 1e-3f); z.w = x.w * inv + (be.w - mu.w * inv);
            st4(a.z0 + (int64_t)gm * F + c, z);
        }
        tile_put4(X0, R, c, r, z);
    });
    __syncthreads();
    float acc[4][4];
    // fc1: F -> 200, relu, dropout
    tile_zero(acc);
    tile_layer<R>(X0, F, a.w1, F1, F1, acc, WB, tid);
    if (cg < F1 / 4) {
        const float4 bias = ld4(a.b1 + 4 * cg);
        const float bj[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gm = row0 + 4 * rg + i;
            // the four columns of a micro-tile row share one Philox block (F1 is a multiple of 4)
            float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (drop) u4 = philox_uniform4(s_lo, s_hi, 1u, step, ((sbase + (uint64_t)gm) * F1 + 4 * cg) >> 2);
            const float uj[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = fmaxf(acc[i][j] + bj[j], 0.f);
                if (drop) v = (uj[j] < keep) ? v / keep : 0.f;
                acc[i][j] = v;
            }
            if (gm < a.B) st4(a.g1 + (int64_t)gm * F1 + 4 * cg, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        }
        tile_store_smem<R>(X1, acc, rg, cg);
    }
    __syncthreads();
    // fc2: 200 -> 80, relu, dropout
    tile_zero(acc);
    tile_layer<R>(X1, F1, a.w2, F2, F2, acc, WB, tid);
    if (cg < F2 / 4) {
        const float4 bias = ld4(a.b2 + 4 * cg);
        const float bj[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gm = row0 + 4 * rg + i;
            // the four columns of a micro-tile row share one Philox block (F2 is a multiple of 4)
            float4 u4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (drop) u4 = philox_uniform4(s_lo, s_hi, 2u, step, ((sbase + (uint64_t)gm) * F2 + 4 * cg) >> 2);
            const float uj[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = fmaxf(acc[i][j] + bj[j], 0.f);
                if (drop) v = (uj[j] < keep) ? v / keep : 0.f;
                acc[i][j] = v;
            }
            if (gm < a.B) st4(a.g2 + (int64_t)gm * F2 + 4 * cg, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        }
        tile_store_smem<R>(X2, acc, rg, cg);
    }
    __syncthreads();
    // fc3 + sigmoid + log-loss (eps 1e-7, score.py:80): one warp per two rows
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < R; r += TL_CT / 32) {
        const int gm = row0 + r;
        float p = 0.f;
        for (int c = lane; c < F2; c += 32) p += X2[c * R + r] * a.w3[c];
        p = warp_sum(p) + a.b3[0];
        if (lane == 0 && gm < a.B) {
            const float pr = sigmoidf_acc(p);
            const float yl = (float)a.label[gm];
            const float eps = 1e-7f;
            a.y[gm] = pr;
            a.loss_b[gm] = -yl * logf(pr + eps) - (1.f - yl) * logf(1.f - pr + eps);
            const float dp = (-yl / (pr + eps) + (1.f - yl) / (1.f - pr + eps)) * a.hp->inv_batch;
            a.dlogit[gm] = dp * pr * (1.f - pr);
        }
    }
}

void launch_fc_fwd(cudaStream_t st, const FcArgs& a) {
    const size_t smem = ((size_t)(a.F + F1 + F2) * R + TileGeom<R>::WBUF) * sizeof(float);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaFuncSetAttribute(fc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = smem;
    }
    launch_chain(fc_fwd_kernel, dim3((a.B + R - 1) / R), dim3(TL_CT), smem, st, a);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------ backward
// dg2 = dlogit w3^T (*) mask(g2);  dg1 = (dg2 W2^T) (*) mask(g1);  dz0 = dg1 W1^T;  dfc_in = dz0 * gamma/sqrt(var+eps)
__global__ void __launch_bounds__(TL_CT) fc_bwd_kernel(FcBwdArgs a) {
    pdl_enter();
    using G = TileGeom<R>;
    extern __shared__ __align__(16) float sm[];
    const int F = a.F, tid = threadIdx.x;
    float* D2 = sm;                 // [80][R]
    float* D1 = D2 + F2 * R;        // [200][R]
    float* WB = D1 + F1 * R;        // weight staging
    const int row0 = blockIdx.x * R;
    const float keep = a.hp->keep_prob;
    const bool drop = (a.hp->train != 0) && keep < 1.f;
    int rg, cg;
    tile_coords<R>(tid, rg, cg);
    tile_for_each_chunk<R>(F2, tid, [&](int r, int c4) {
        const int gm = row0 + r, c = 4 * c4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gm < a.B) {
            const float4 g = ld4(a.g2 + (int64_t)gm * F2 + c), w = ld4(a.w3 + c);
            const float dl = a.dlogit[gm];
            float t;
            t = dl * w.x; v.x = (g.x > 0.f) ? (drop ? t / keep : t) : 0.f;
            t = dl * w.y; v.y = (g.y > 0.f) ? (drop ? t / keep : t) : 0.f;
            t = dl * w.z; v.z = (g.z > 0.f) ? (drop ? t / keep : t) : 0.f;
            t = dl * w.w; v.w = (g.w > 0.f) ? (drop ? t / keep : t) : 0.f;
            st4(a.dg2 + (int64_t)gm * F2 + c, v);
        }
        tile_put4(D2, R, c, r, v);
    });
    __syncthreads();
    float acc[4][4];
    tile_zero(acc);
    tile_layer<R>(D2, F2, a.w2t, F1, F1, acc, WB, tid);
    if (cg < F1 / 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int gm = row0 + 4 * rg + i;
            if (gm < a.B) {
                const float4 g = ld4(a.g1 + (int64_t)gm * F1 + 4 * cg);
                acc[i][0] = (g.x > 0.f) ? (drop ? acc[i][0] / keep : acc[i][0]) : 0.f;
                acc[i][1] = (g.y > 0.f) ? (drop ? acc[i][1] / keep : acc[i][1]) : 0.f;
                acc[i][2] = (g.z > 0.f) ? (drop ? acc[i][2] / keep : acc[i][2]) : 0.f;
                acc[i][3] = (g.w > 0.f) ? (drop ? acc[i][3] / keep : acc[i][3]) : 0.f;
                st4(a.dg1 + (int64_t)gm * F1 + 4 * cg, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
            } else {
                acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            }
        }
        tile_store_smem<R>(D1, acc, rg, cg);
    }
    __syncthreads();
    for (int cb = 0; cb < F; cb += G::NP) {
        const int nb = min(G::NP, F - cb);
        tile_zero(acc);
        tile_layer<R>(D1, F1, a.w1t + cb, F, nb, acc, WB, tid);
        if (4 * cg < nb) {
            const int c = cb + 4 * cg;
            const float4 ga = ld4(a.gamma + c), va = ld4(a.var + c);
            const float inv[4] = {ga.x / sqrtf(va.x + 1e-3f), ga.y / sqrtf(va.y + 1e-3f), ga.z / sqrtf(va.z + 1e-3f),
                                  ga.w / sqrtf(va.w + 1e-3f)};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int gm = row0 + 4 * rg + i;
                if (gm < a.B) {
                    st4(a.dz0 + (int64_t)gm * F + c, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
                    st4(a.dfc_in + (int64_t)gm * F + c,
                        make_float4(acc[i][0] * inv[0], acc[i][1] * inv[1], acc[i][2] * inv[2], acc[i][3] * inv[3]));
                }
            }
        }
    }
}

void launch_fc_bwd(cudaStream_t st, const FcBwdArgs& a) {
    const size_t smem = ((size_t)(F1 + F2) * R + TileGeom<R>::WBUF) * sizeof(float);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaFuncSetAttribute(fc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = smem;
    }
    launch_chain(fc_bwd_kernel, dim3((a.B + R - 1) / R), dim3(TL_CT), smem, st, a);
    ++g_launch_count;
}

}  // namespace score
