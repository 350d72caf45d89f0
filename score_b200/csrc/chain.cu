// Row-tile fused MLP chains.
//
// The prediction head of the path (build_fc_net + log_loss, score.py:68-81) is a chain of tiny layers on B rows:
// BN -> 200 relu dropout -> 80 relu dropout -> 1 -> sigmoid -> loss.  As separate SGEMM launches each layer is pure
// launch + pipeline-fill latency (ncu launch list, profiles/): here one CTA carries a tile of R = 16 rows through the
// whole chain with the activations in shared memory, stored k-major ([k][R]) so a thread that owns one output
// column reads four rows per LDS.128 (a warp-wide broadcast) and streams its weight column from L2.
// fc_bwd_kernel is the same idea for the backward chain; the weight gradients stay SGEMMs on the side stream.
#include "kernels.h"

namespace score {

namespace {

constexpr int R = 16;          // rows per CTA
constexpr int CT = 256;        // threads per CTA

constexpr int KC = 16;         // weight rows staged per pipeline stage
constexpr int NST = 4;         // cp.async pipeline depth of the weight stream

__device__ __forceinline__ void cpa4(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}

// xk points at the 16 rows this thread owns inside one k-row of the activation tile
__device__ __forceinline__ void fma_rows(const float* __restrict__ xk, float w, float (&acc)[R]) {
    const float4* x = reinterpret_cast<const float4*>(xk);
#pragma unroll
    for (int q = 0; q < R / 4; ++q) {
        const float4 v = x[q];
        acc[4 * q + 0] = fmaf(v.x, w, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(v.y, w, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(v.z, w, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(v.w, w, acc[4 * q + 3]);
    }
}

// One layer of a row-tile chain, called by ALL threads of the CTA:
//   acc[r] (+)= sum_k Xs[k][r] * Wop(k, n)    for the column n this thread owns (n < 0: the thread only helps staging)
// TRANS = false: Wop(k,n) = W[k*ldw + n]  (forward use of a row-major [K][N] kernel)
// TRANS = true : Wop(k,n) = W[n*ldw + k]  (backward use: dX = dY W^T with the same storage)
// The weights are streamed through shared memory in chunks of KC rows with a 2-stage cp.async pipeline so the
// inner loop never waits on a global load.  wbuf: chain_wbuf_floats(N) floats.
// XR: floats per k-row of the activation tile (R, or 2R when the CTA carries two 16-row halves).
template <bool TRANS, int XR = R>
__device__ __forceinline__ void chain_layer(const float* __restrict__ Xs, int K, const float* __restrict__ W, int ldw,
                                            int N, int n, float (&acc)[R], float* wbuf, int tid) {
    const int wstride = TRANS ? KC + 1 : ((N + 3) & ~3);
    const int stage_floats = TRANS ? N * wstride : KC * wstride;
    const bool vec = !TRANS && ((ldw & 3) == 0) && ((N & 3) == 0) && ((((uintptr_t)W) & 15) == 0);
    auto stage = [&](int c) {
        float* dst = wbuf + (c % NST) * stage_floats;
        const int k0 = c * KC;
        if (k0 < K) {
            if (TRANS) {
                for (int i = tid; i < N * KC; i += CT) {
                    const int nn = i / KC, k = i - nn * KC;
                    const bool ok = k0 + k < K;
                    cpa4(dst + nn * wstride + k, ok ? W + (int64_t)nn * ldw + k0 + k : W, ok ? 4 : 0);
                }
            } else if (vec) {
                const int n4 = N >> 2;
                for (int i = tid; i < KC * n4; i += CT) {
                    const int k = i / n4, c4 = i - k * n4;
                    const bool ok = k0 + k < K;
                    cp_async16(dst + k * wstride + c4 * 4, ok ? W + (int64_t)(k0 + k) * ldw + c4 * 4 : W, ok ? 16 : 0);
                }
            } else {
                for (int i = tid; i < KC * N; i += CT) {
                    const int k = i / N, nn = i - k * N;
                    const bool ok = k0 + k < K;
                    cpa4(dst + k * wstride + nn, ok ? W + (int64_t)(k0 + k) * ldw + nn : W, ok ? 4 : 0);
                }
            }
        }
        cp_async_commit();
    };
    const int nchunks = (K + KC - 1) / KC;
#pragma unroll
    for (int c = 0; c < NST - 1; ++c) stage(c);
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<NST - 2>();
        __syncthreads();
        stage(c + NST - 1);   // refills the slot read in iteration c-1 (every thread is past the barrier)
        if (n >= 0) {
            const float* wb = wbuf + (c % NST) * stage_floats;
            const int kmax = min(KC, K - c * KC);
#pragma unroll 4
            for (int k = 0; k < kmax; ++k) {
                const float w = TRANS ? wb[n * wstride + k] : wb[k * wstride + n];
                fma_rows(Xs + (c * KC + k) * XR, w, acc);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();   // the caller may restage wbuf / overwrite Xs right away
}
// floats of weight staging a layer with N outputs needs
__host__ __device__ inline int chain_wbuf_floats(int N) { return NST * max(KC * ((N + 3) & ~3), N * (KC + 1)); }

}  // namespace

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(CT) fc_fwd_kernel(FcArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int F = a.F, tid = threadIdx.x;
    float* X0 = sm;                 // [F][R]   bn1 output
    float* X1 = X0 + F * R;         // [200][R] fc1 output (after dropout)
    float* X2 = X1 + 200 * R;       // [80][R]  fc2 output (after dropout)
    float* WB = X2 + 80 * R;        // weight staging
    const int row0 = blockIdx.x * R;
    const float keep = a.hp->keep_prob;
    const bool drop = (a.hp->train != 0) && keep < 1.f;
    const uint32_t s_lo = a.hp->seed_lo, s_hi = a.hp->seed_hi, step = (uint32_t)a.hp->step;
    // batch norm in inference mode (score.py:69): consecutive threads take consecutive rows of one column
    for (int i = tid; i < F * R; i += CT) {
        const int r = i % R, c = i / R, gm = row0 + r;
        float z = 0.f;
        if (gm < a.B) {
            const float inv = a.gamma[c] / sqrtf(a.var[c] + 1e-3f);
            z = a.fc_in[(int64_t)gm * F + c] * inv + (a.beta[c] - a.mean[c] * inv);
            a.z0[(int64_t)gm * F + c] = z;
        }
        X0[c * R + r] = z;
    }
    __syncthreads();
    // fc1: F -> 200, relu, dropout
    {
        const int n = tid < 200 ? tid : -1;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        chain_layer<false>(X0, F, a.w1, 200, 200, n, acc, WB, tid);
        const float bias = n >= 0 ? a.b1[n] : 0.f;
        if (n >= 0)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int gm = row0 + r;
            float v = fmaxf(acc[r] + bias, 0.f);
            if (drop) {
                const float u = philox_uniform(s_lo, s_hi, 1u, step, (uint64_t)gm * 200 + n);
                v = (u < keep) ? v / keep : 0.f;
            }
            X1[n * R + r] = v;
            if (gm < a.B) a.g1[(int64_t)gm * 200 + n] = v;
        }
    }
    __syncthreads();
    // fc2: 200 -> 80, relu, dropout
    {
        const int n = tid < 80 ? tid : -1;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        chain_layer<false>(X1, 200, a.w2, 80, 80, n, acc, WB, tid);
        const float bias = n >= 0 ? a.b2[n] : 0.f;
        if (n >= 0)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int gm = row0 + r;
            float v = fmaxf(acc[r] + bias, 0.f);
            if (drop) {
                const float u = philox_uniform(s_lo, s_hi, 2u, step, (uint64_t)gm * 80 + n);
                v = (u < keep) ? v / keep : 0.f;
            }
            X2[n * R + r] = v;
            if (gm < a.B) a.g2[(int64_t)gm * 80 + n] = v;
        }
    }
    __syncthreads();
    // fc3 + sigmoid + log-loss (eps 1e-7, score.py:80): one warp per two rows
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < R; r += CT / 32) {
        const int gm = row0 + r;
        float p = 0.f;
        for (int c = lane; c < 80; c += 32) p += X2[c * R + r] * a.w3[c];
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
    const size_t smem = ((size_t)(a.F + 200 + 80) * R + chain_wbuf_floats(200)) * sizeof(float);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaFuncSetAttribute(fc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = smem;
    }
    fc_fwd_kernel<<<(a.B + R - 1) / R, CT, smem, st>>>(a);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------ backward
// dg2 = dlogit w3^T (*) mask(g2);  dg1 = (dg2 W2^T) (*) mask(g1);  dz0 = dg1 W1^T;  dfc_in = dz0 * gamma/sqrt(var+eps)
__global__ void __launch_bounds__(CT) fc_bwd_kernel(FcBwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int F = a.F, tid = threadIdx.x;
    float* D2 = sm;                 // [80][R]
    float* D1 = D2 + 80 * R;        // [200][R]
    float* WB = D1 + 200 * R;       // weight staging
    const int row0 = blockIdx.x * R;
    const float keep = a.hp->keep_prob;
    const bool drop = (a.hp->train != 0) && keep < 1.f;
    for (int i = tid; i < 80 * R; i += CT) {
        const int r = i % R, j = i / R, gm = row0 + r;
        float v = 0.f;
        if (gm < a.B) {
            const float g = a.g2[(int64_t)gm * 80 + j];
            v = a.dlogit[gm] * a.w3[j];
            v = (g > 0.f) ? (drop ? v / keep : v) : 0.f;
            a.dg2[(int64_t)gm * 80 + j] = v;
        }
        D2[j * R + r] = v;
    }
    __syncthreads();
    {
        const int n = tid < 200 ? tid : -1;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        chain_layer<true>(D2, 80, a.w2, 80, 200, n, acc, WB, tid);
        if (n >= 0)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int gm = row0 + r;
            float v = 0.f;
            if (gm < a.B) {
                const float g = a.g1[(int64_t)gm * 200 + n];
                v = (g > 0.f) ? (drop ? acc[r] / keep : acc[r]) : 0.f;
                a.dg1[(int64_t)gm * 200 + n] = v;
            }
            D1[n * R + r] = v;
        }
    }
    __syncthreads();
    for (int nb = 0; nb < F; nb += CT) {   // F may exceed the CTA width: column blocks, every thread helps staging
        const int n = (nb + tid < F) ? nb + tid : -1;
        const int ncols = min(CT, F - nb);
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        chain_layer<true>(D1, 200, a.w1 + (int64_t)nb * 200, 200, ncols, n >= 0 ? n - nb : -1, acc, WB, tid);
        const float inv = n >= 0 ? a.gamma[n] / sqrtf(a.var[n] + 1e-3f) : 0.f;
        if (n >= 0)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int gm = row0 + r;
            if (gm < a.B) {
                a.dz0[(int64_t)gm * F + n] = acc[r];
                a.dfc_in[(int64_t)gm * F + n] = acc[r] * inv;
            }
        }
    }
}

void launch_fc_bwd(cudaStream_t st, const FcBwdArgs& a) {
    const size_t smem = ((size_t)(200 + 80) * R + chain_wbuf_floats(CT)) * sizeof(float);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaFuncSetAttribute(fc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = smem;
    }
    fc_bwd_kernel<<<(a.B + R - 1) / R, CT, smem, st>>>(a);
    ++g_launch_count;
}

// ------------------------------------------------------------------------------------------ attention chain
// attention() of score.py:169-186 for a tile of 32 (b,t) rows per CTA, two 16-row halves: threads [0,128) carry
// rows 0-15, threads [128,256) rows 16-31, thread (tid & 127) = output column.
//   inp = [q | key | q-key | q*key]  (written out once: the weight-gradient GEMM of the first layer reads it)
//   f1 = relu(inp W1 + b1) (80)   f2 = relu(f1 W2 + b2) (40)   s = f2 w3 + b3   (masking + softmax: att_pool)
constexpr int AR = 2 * R;   // rows per CTA

__global__ void __launch_bounds__(CT) att_fwd_kernel(AttChainArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int Dk = a.Dk, K1 = 4 * Dk, tid = threadIdx.x;
    float* X0 = sm;                    // [4Dk][AR]
    float* X1 = X0 + K1 * AR;          // [80][AR]
    float* X2 = X1 + 80 * AR;          // [40][AR]
    float* WB = X2 + 40 * AR;
    const int64_t row0 = (int64_t)blockIdx.x * AR;
    for (int i = tid; i < Dk * AR; i += CT) {
        const int r = i % AR, c = i / AR;
        const int64_t m = row0 + r;
        float qv = 0.f, kv = 0.f;
        if (m < a.M) {
            qv = a.q[(int64_t)((int)m / a.T) * Dk + c];
            kv = a.key[m * Dk + c];
            float* o = a.a1 + m * K1;
            o[c] = qv; o[Dk + c] = kv; o[2 * Dk + c] = qv - kv; o[3 * Dk + c] = qv * kv;
        }
        X0[c * AR + r] = qv; X0[(Dk + c) * AR + r] = kv;
        X0[(2 * Dk + c) * AR + r] = qv - kv; X0[(3 * Dk + c) * AR + r] = qv * kv;
    }
    __syncthreads();
    const int half = tid >> 7, col = tid & 127;
    {
        const int n = col < 80 ? col : -1;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        chain_layer<false, AR>(X0 + half * R, K1, a.w1, 80, 80, n, acc, WB, tid);
        if (n >= 0) {
            const float bias = a.b1[n];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int64_t m = row0 + half * R + r;
                const float v = fmaxf(acc[r] + bias, 0.f);
                X1[n * AR + half * R + r] = v;
                if (m < a.M) a.f1[m * 80 + n] = v;
            }
        }
    }
    __syncthreads();
    {
        const int n = col < 40 ? col : -1;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        chain_layer<false, AR>(X1 + half * R, 80, a.w2, 40, 40, n, acc, WB, tid);
        if (n >= 0) {
            const float bias = a.b2[n];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int64_t m = row0 + half * R + r;
                const float v = fmaxf(acc[r] + bias, 0.f);
                X2[n * AR + half * R + r] = v;
                if (m < a.M) a.f2[m * 40 + n] = v;
            }
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < AR; r += CT / 32) {
        const int64_t m = row0 + r;
        float p = 0.f;
        for (int c = lane; c < 40; c += 32) p += X2[c * AR + r] * a.w3[c];
        p = warp_sum(p) + a.b3[0];
        if (lane == 0 && m < a.M) a.s[m] = p;
    }
}

void launch_att_fwd(cudaStream_t st, const AttChainArgs& a) {
    const size_t smem = ((size_t)(4 * a.Dk + 80 + 40) * AR + chain_wbuf_floats(80)) * sizeof(float);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaFuncSetAttribute(att_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = smem;
    }
    att_fwd_kernel<<<(unsigned)((a.M + AR - 1) / AR), CT, smem, st>>>(a);
    ++g_launch_count;
}

// backward data chain of attention(): ds -> df2 -> df1 -> d inp (never written) -> dkey, per-row dq.
//   d inp = [dA | dB | dC | dD]:   dq_row = dA + dC + dD*key     dkey = dB - dC + dD*q  (+ pooling gradient for c < acc_cols)
// Thread (half, c) owns key column c for 16 rows; requires Dk <= 128.
__global__ void __launch_bounds__(CT) att_bwd_kernel(AttChainBwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int Dk = a.Dk, tid = threadIdx.x;
    float* D2 = sm;                    // [40][AR]
    float* D1 = D2 + 40 * AR;          // [80][AR]
    float* Qs = D1 + 80 * AR;          // [Dk][AR]
    float* Ks = Qs + Dk * AR;          // [Dk][AR]
    float* WB = Ks + Dk * AR;
    const int64_t row0 = (int64_t)blockIdx.x * AR;
    for (int i = tid; i < Dk * AR; i += CT) {
        const int r = i % AR, c = i / AR;
        const int64_t m = row0 + r;
        float qv = 0.f, kv = 0.f;
        if (m < a.M) { qv = a.q[(int64_t)((int)m / a.T) * Dk + c]; kv = a.key[m * Dk + c]; }
        Qs[c * AR + r] = qv; Ks[c * AR + r] = kv;
    }
    for (int i = tid; i < 40 * AR; i += CT) {
        const int r = i % AR, j = i / AR;
        const int64_t m = row0 + r;
        float v = 0.f;
        if (m < a.M) {
            v = (a.f2[m * 40 + j] > 0.f) ? a.ds[m] * a.w3[j] : 0.f;
            a.df2[m * 40 + j] = v;
        }
        D2[j * AR + r] = v;
    }
    __syncthreads();
    const int half = tid >> 7, col = tid & 127;
    {
        const int n = col < 80 ? col : -1;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        chain_layer<true, AR>(D2 + half * R, 40, a.w2, 40, 80, n, acc, WB, tid);
        if (n >= 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int64_t m = row0 + half * R + r;
                float v = 0.f;
                if (m < a.M) {
                    v = (a.f1[m * 80 + n] > 0.f) ? acc[r] : 0.f;
                    a.df1[m * 80 + n] = v;
                }
                D1[n * AR + half * R + r] = v;
            }
        }
    }
    __syncthreads();
    const int c = col < Dk ? col : -1;
    float dq[R], dk[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { dq[r] = 0.f; dk[r] = 0.f; }
    for (int part = 0; part < 4; ++part) {
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        chain_layer<true, AR>(D1 + half * R, 80, a.w1 + (int64_t)part * Dk * 80, 80, Dk, c, acc, WB, tid);
        if (c >= 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float qv = Qs[c * AR + half * R + r], kv = Ks[c * AR + half * R + r];
                if (part == 0) dq[r] += acc[r];
                else if (part == 1) dk[r] += acc[r];
                else if (part == 2) { dq[r] += acc[r]; dk[r] -= acc[r]; }
                else { dq[r] += acc[r] * kv; dk[r] += acc[r] * qv; }
            }
        }
    }
    if (c >= 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t m = row0 + half * R + r;
            if (m < a.M) {
                float v = dk[r];
                if (c < a.acc_cols) v += a.dkey[m * Dk + c];
                a.dkey[m * Dk + c] = v;
                a.dq_row[m * Dk + c] = dq[r];
            }
        }
    }
}

void launch_att_bwd(cudaStream_t st, const AttChainBwdArgs& a) {
    const size_t smem = ((size_t)(40 + 80 + 2 * a.Dk) * AR + chain_wbuf_floats(a.Dk > 80 ? a.Dk : 80)) * sizeof(float);
    static size_t attr_set = 0;
    if (smem > 48 * 1024 && smem > attr_set) {
        cudaFuncSetAttribute(att_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = smem;
    }
    att_bwd_kernel<<<(unsigned)((a.M + AR - 1) / AR), CT, smem, st>>>(a);
    ++g_launch_count;
}

// dq[b][c] = sum_t dq_row[b*T + t][c]   (fixed order)
__global__ void dq_reduce_kernel(int B, int T, int Dk, const float* __restrict__ dq_row, float* __restrict__ dq) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * Dk) return;
    const int b = idx / Dk, c = idx - b * Dk;
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += dq_row[((int64_t)b * T + t) * Dk + c];
    dq[idx] = s;
}
void launch_dq_reduce(cudaStream_t st, int B, int T, int Dk, const float* dq_row, float* dq) {
    dq_reduce_kernel<<<(B * Dk + 255) / 256, 256, 0, st>>>(B, T, Dk, dq_row, dq);
    ++g_launch_count;
}

}  // namespace score
