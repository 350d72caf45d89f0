// Dense layers of the SCoRe path: one fp32 SGEMM with fused epilogues.
//
// Every tf.layers.dense on the path (score.py:69-74,156,172-177), the GRU input projection
// (score.py:205-208) and all their backward contractions are tall-skinny fp32 products
// (M = B*T or B rows, N <= 416, K <= 416).  fp32 FFMA is used on purpose: the parity bar is
// 1e-5 relative in fp32 (BASELINE.json), which tf32/bf16 tensor-core inputs cannot meet, and
// the whole dense block is ~1-15 MFLOP/sample.
//
// Operands are addressed by (row stride, col stride) so the same kernel serves
//   forward          C = act(A W + b)
//   backward data    dA = (dC W^T) (*) relu/dropout mask
//   backward weight  dW = A^T dC, split over the batch rows into fixed partials (+ bias colsum)
// Split partials are reduced in a fixed order by reduce_partials -> deterministic gradients.
#include "kernels.h"

namespace score {

int64_t g_launch_count = 0;

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_kernel(GemmArgs p) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int LA = BM * BK / NT;   // A elements per thread per tile
    constexpr int LB = BN * BK / NT;
    static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/thread mismatch");
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    // K range of this split
    int k_begin = 0, k_end = p.K;
    if (p.splits > 1) {
        int chunk = (p.K + p.splits - 1) / p.splits;
        chunk = (chunk + BK - 1) / BK * BK;
        k_begin = blockIdx.z * chunk;
        k_end = min(p.K, k_begin + chunk);
    }
    const bool a_kcontig = (p.a_cs == 1);
    const bool b_kcontig = (p.b_rs == 1 && p.b_cs != 1);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float csum[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) csum[j] = 0.f;
    const bool do_colsum = (p.colsum != nullptr) && blockIdx.x == 0 && ty == 0;

    float ra[LA], rb[LB];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            int idx = tid + i * NT;
            int m, k;
            if (a_kcontig) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
            int gm = m0 + m, gk = k0 + k;
            ra[i] = (gm < p.M && gk < k_end) ? __ldg(p.A + (int64_t)gm * p.a_rs + (int64_t)gk * p.a_cs) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            int idx = tid + i * NT;
            int n, k;
            if (b_kcontig) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
            int gn = n0 + n, gk = k0 + k;
            rb[i] = (gn < p.N && gk < k_end) ? __ldg(p.B + (int64_t)gk * p.b_rs + (int64_t)gn * p.b_cs) : 0.f;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            int idx = tid + i * NT;
            int m, k;
            if (a_kcontig) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
            As[k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            int idx = tid + i * NT;
            int n, k;
            if (b_kcontig) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
            Bs[k][n] = rb[i];
        }
    };

    if (k_begin < k_end) {
        load_tiles(k_begin);
        store_tiles();
        __syncthreads();
        for (int k0 = k_begin; k0 < k_end; k0 += BK) {
            const bool more = (k0 + BK) < k_end;
            if (more) load_tiles(k0 + BK);
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float a[TM], b[TN];
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    float4 t = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
                    a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
                }
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    float4 t = *reinterpret_cast<const float4*>(&Bs[k][tx * TN + j]);
                    b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                if (do_colsum) {
#pragma unroll
                    for (int j = 0; j < TN; ++j) csum[j] += b[j];
                }
            }
            __syncthreads();
            if (more) {
                store_tiles();
                __syncthreads();
            }
        }
    }

    // ---- epilogue
    float* C = p.C;
    if (p.epi == EPI_SPLIT) C += (int64_t)blockIdx.z * p.c_split_stride;
    float keep = 1.f;
    bool drop = false;
    uint32_t s_lo = 0, s_hi = 0, step = 0;
    if (p.epi == EPI_BIAS_RELU_DROP || (p.epi == EPI_MASK && p.mask_dropout)) {
        keep = p.hp->keep_prob;
        drop = (p.hp->train != 0) && keep < 1.f;
        s_lo = p.hp->seed_lo; s_hi = p.hp->seed_hi; step = (uint32_t)p.hp->step;
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int gm = m0 + ty * TM + i;
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int gn = n0 + tx * TN + j;
            if (gn >= p.N) continue;
            float v = acc[i][j];
            float* dst = C + (int64_t)gm * p.c_rs + gn;
            switch (p.epi) {
                case EPI_BIAS: v += p.bias[gn]; break;
                case EPI_BIAS_RELU: v = fmaxf(v + p.bias[gn], 0.f); break;
                case EPI_BIAS_RELU_DROP: {
                    v = fmaxf(v + p.bias[gn], 0.f);
                    if (drop) {
                        float u = philox_uniform(s_lo, s_hi, p.rng_stream, step, (uint64_t)gm * p.N + gn);
                        v = (u < keep) ? v / keep : 0.f;
                    }
                    break;
                }
                case EPI_MASK: {
                    float a = p.aux[(int64_t)gm * p.aux_rs + gn];
                    v = (a > 0.f) ? (drop ? v / keep : v) : 0.f;
                    break;
                }
                case EPI_ACCUM: v += *dst; break;
                default: break;
            }
            *dst = v;
        }
    }
    if (do_colsum) {
        float* cs = p.colsum + (int64_t)blockIdx.z * p.colsum_split_stride;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int gn = n0 + tx * TN + j;
            if (gn < p.N) cs[gn] = csum[j];
        }
    }
}

void launch_gemm(cudaStream_t st, const GemmArgs& a) {
    if (a.M <= 0 || a.N <= 0) return;
    constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
    dim3 grid((a.M + BM - 1) / BM, (a.N + BN - 1) / BN, a.splits > 1 ? a.splits : 1);
    gemm_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(a);
    ++g_launch_count;
}

}  // namespace score
