// Dense layers of the SCoRe path: one fp32 SGEMM family with fused epilogues.
//
// Every tf.layers.dense on the path (score.py:69-74,156,172-177), the GRU input projection
// (score.py:205-208) and all their backward contractions are tall-skinny fp32 products
// (M = B*T or B rows, N <= 416, K <= 416).  fp32 FFMA is used on purpose: the parity bar is
// 1e-5 relative in fp32 (BASELINE.json), which tf32/bf16 tensor-core inputs cannot meet, and
// the whole dense block is ~1-15 MFLOP/sample.
//
// Operands are addressed by (row stride, col stride) so the same kernel serves
//   forward          C = act(A W + b)                         A k-contiguous, W n-contiguous
//   backward data    dA = (dC W^T) (*) relu/dropout mask      dC k-contiguous, W^T k-contiguous
//   backward weight  dW = A^T dC, split over the batch rows   A^T m-contiguous, dC n-contiguous
// Tiles are staged with a 3-stage cp.async pipeline (16-byte copies when the contiguous dimension
// is 16-byte aligned, 4-byte copies otherwise); each operand is kept in shared memory in the
// orientation of its contiguous global dimension so the copies never transpose.
// Split partials are reduced in a fixed order by reduce_partials -> deterministic gradients.
#include <stdlib.h>

#include "kernels.h"

namespace score {

int64_t g_launch_count = 0;
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("SCORE_PDL"); on = (e && atoi(e) != 0) ? 1 : 0; }
    return on == 1;
}

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4, STAGES = 3;
constexpr int NT = (BM / TM) * (BN / TN);   // 256 threads
constexpr int LDK = BK + 4;                 // row stride of a [rows][k] tile (k contiguous)
constexpr int LDR = BM + 4;                 // row stride of a [k][rows] tile (rows contiguous); BM == BN

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async16_ca(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}

// Stage one [64 rows x 16 k] operand tile.  KMAJ: the operand is k-contiguous in global memory and is kept as
// tile[row][k]; otherwise it is row-contiguous and kept as tile[k][row].
//   elem(row, k) = base[row * rs + k * cs]
template <bool KMAJ>
__device__ __forceinline__ void stage_tile(float* tile, const float* __restrict__ base, int64_t rs, int64_t cs,
                                           int row0, int nrows, int k0, int k_end, bool vec, int tid) {
    if (vec) {
        if (KMAJ) {
            const int r = tid >> 2, kq = (tid & 3) * 4;
            const int gr = row0 + r, gk = k0 + kq;
            int bytes = 0;
            if (gr < nrows && gk < k_end) bytes = min(4, k_end - gk) * 4;
            const float* src = bytes ? base + (int64_t)gr * rs + gk : base;
            cp_async16_ca(tile + r * LDK + kq, src, bytes);
        } else {
            const int k = tid >> 4, rq = (tid & 15) * 4;
            const int gr = row0 + rq, gk = k0 + k;
            int bytes = 0;
            if (gk < k_end && gr < nrows) bytes = min(4, nrows - gr) * 4;
            const float* src = bytes ? base + (int64_t)gk * cs + gr : base;
            cp_async16_ca(tile + k * LDR + rq, src, bytes);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * NT;
            int r, k;
            if (KMAJ) { k = idx & (BK - 1); r = idx >> 4; } else { r = idx & (BM - 1); k = idx >> 6; }
            const int gr = row0 + r, gk = k0 + k;
            const bool ok = gr < nrows && gk < k_end;
            const float* src = ok ? base + (int64_t)gr * rs + (int64_t)gk * cs : base;
            cp_async4(KMAJ ? tile + r * LDK + k : tile + k * LDR + r, src, ok ? 4 : 0);
        }
    }
}

template <bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(NT) gemm_kernel(GemmBatch batch) {
    // blockIdx.z = problem * splits + split: independent problems of one launch share the grid
    const int nsplit = batch.g[0].splits > 1 ? batch.g[0].splits : 1;
    const GemmArgs& p = batch.g[blockIdx.z / nsplit];
    const int split = blockIdx.z % nsplit;
    constexpr int A_TILE = A_KMAJ ? BM * LDK : BK * LDR;
    constexpr int B_TILE = B_KMAJ ? BN * LDK : BK * LDR;
    __shared__ __align__(16) float As[STAGES][A_TILE];
    __shared__ __align__(16) float Bs[STAGES][B_TILE];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (m0 >= p.M || n0 >= p.N) return;   // the grid is sized for the largest problem of the batch

    int k_begin = 0, k_end = p.K;
    if (p.splits > 1) {
        int chunk = (p.K + p.splits - 1) / p.splits;
        chunk = (chunk + BK - 1) / BK * BK;
        k_begin = split * chunk;
        k_end = min(p.K, k_begin + chunk);
    }
    // 16-byte copies need the contiguous dimension's base and leading stride 16-byte aligned
    const bool a_vec = A_KMAJ ? (((p.a_rs & 3) == 0) && ((((uintptr_t)p.A) & 15) == 0) && ((k_begin & 3) == 0))
                              : (((p.a_cs & 3) == 0) && ((((uintptr_t)p.A) & 15) == 0));
    const bool b_vec = B_KMAJ ? (((p.b_cs & 3) == 0) && ((((uintptr_t)p.B) & 15) == 0) && ((k_begin & 3) == 0))
                              : (((p.b_rs & 3) == 0) && ((((uintptr_t)p.B) & 15) == 0));

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float csum[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) csum[j] = 0.f;
    const bool do_colsum = (p.colsum != nullptr) && blockIdx.x == 0 && ty == 0;

    const int nk = (k_end > k_begin) ? (k_end - k_begin + BK - 1) / BK : 0;
    auto issue = [&](int kt) {
        if (kt < nk) {
            const int s = kt % STAGES, k0 = k_begin + kt * BK;
            // A: rows = m, elem(m,k) = A[m*a_rs + k*a_cs];  B: rows = n, elem(n,k) = B[k*b_rs + n*b_cs]
            stage_tile<A_KMAJ>(As[s], p.A, p.a_rs, p.a_cs, m0, p.M, k0, k_end, a_vec, tid);
            stage_tile<B_KMAJ>(Bs[s], p.B, p.b_cs, p.b_rs, n0, p.N, k0, k_end, b_vec, tid);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);

    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        issue(kt + STAGES - 1);   // refills the stage consumed in iteration kt-1 (all threads passed the barrier)
        const float* a_t = As[kt % STAGES];
        const float* b_t = Bs[kt % STAGES];
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            float a[4][TM], b[4][TN];   // [k][row]
            if (A_KMAJ) {
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    float4 t = *reinterpret_cast<const float4*>(a_t + (ty * TM + i) * LDK + kk);
                    a[0][i] = t.x; a[1][i] = t.y; a[2][i] = t.z; a[3][i] = t.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float4 t = *reinterpret_cast<const float4*>(a_t + (kk + k) * LDR + ty * TM);
                    a[k][0] = t.x; a[k][1] = t.y; a[k][2] = t.z; a[k][3] = t.w;
                }
            }
            if (B_KMAJ) {
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    float4 t = *reinterpret_cast<const float4*>(b_t + (tx * TN + j) * LDK + kk);
                    b[0][j] = t.x; b[1][j] = t.y; b[2][j] = t.z; b[3][j] = t.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float4 t = *reinterpret_cast<const float4*>(b_t + (kk + k) * LDR + tx * TN);
                    b[k][0] = t.x; b[k][1] = t.y; b[k][2] = t.z; b[k][3] = t.w;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[k][i], b[k][j], acc[i][j]);
                if (do_colsum) {
#pragma unroll
                    for (int j = 0; j < TN; ++j) csum[j] += b[k][j];
                }
            }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue
    float* C = p.C;
    if (p.epi == EPI_SPLIT) C += (int64_t)split * p.c_split_stride;
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
        const int gm = m0 + ty * TM + i;
        if (gm >= p.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gn = n0 + tx * TN + j;
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
        float* cs = p.colsum + (int64_t)split * p.colsum_split_stride;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gn = n0 + tx * TN + j;
            if (gn < p.N) cs[gn] = csum[j];
        }
    }
}

}  // namespace

// Independent problems with the same operand orientation and split count in ONE launch.
void launch_gemm_batch(cudaStream_t st, const GemmArgs* list, int n) {
    if (n <= 0) return;
    GemmBatch b{};
    int gm = 0, gn = 0;
    const int nsplit = list[0].splits > 1 ? list[0].splits : 1;
    for (int i = 0; i < n && i < 8; ++i) {
        b.g[i] = list[i];
        gm = max(gm, (list[i].M + BM - 1) / BM);
        gn = max(gn, (list[i].N + BN - 1) / BN);
    }
    if (gm == 0 || gn == 0) return;
    dim3 grid(gm, gn, n * nsplit);
    // orientation of each operand = its contiguous global dimension (k-contiguous wins a tie)
    const bool a_kmaj = (list[0].a_cs == 1);
    const bool b_kmaj = (list[0].b_rs == 1) && (list[0].b_cs != 1);
    if (a_kmaj && b_kmaj) gemm_kernel<true, true><<<grid, NT, 0, st>>>(b);
    else if (a_kmaj && !b_kmaj) gemm_kernel<true, false><<<grid, NT, 0, st>>>(b);
    else if (!a_kmaj && b_kmaj) gemm_kernel<false, true><<<grid, NT, 0, st>>>(b);
    else gemm_kernel<false, false><<<grid, NT, 0, st>>>(b);
    ++g_launch_count;
}

void launch_gemm(cudaStream_t st, const GemmArgs& a) {
    if (a.M <= 0 || a.N <= 0) return;
    launch_gemm_batch(st, &a, 1);
}

}  // namespace score
