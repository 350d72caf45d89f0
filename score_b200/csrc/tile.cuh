// Row-tile layer primitive shared by the fused dense chains (attention MLP, prediction head, GRU projections).
//
// Every dense layer on the SCoRe path is tall-skinny: a few thousand (b,t) rows against a weight matrix of at most
// a few hundred columns (score.py:69-74, 172-177, 205-208).  A CTA therefore owns a tile of RT rows for a whole
// CHAIN of layers: the activations stay in shared memory, k-major ([k][RT], rows contiguous), and only the weights
// stream through a cp.async ring.  Inside a layer every thread owns a 4x4 register micro-tile, so the inner loop is
// two LDS.128 (four rows of X, four columns of W) per 16 FFMA - the classic SGEMM ratio instead of the one LDS per
// FFMA of a thread-per-column chain.  fp32 FFMA on purpose: the parity bar is 1e-5 relative in fp32.
#pragma once
#include "common.cuh"

namespace score {

constexpr int TL_CT = 256;    // threads per CTA (8 warps)
// compile-time tuning points (tools/build_variants.py builds the alternatives for an A/B on the GPU box).  Measured on
// B200, Taobao step: (KC, NST) = (16,4) 0.389 ms, (16,8) 0.372, (32,4) 0.377, (32,6) 0.407 - the chains wait on the
// weight stream, so a deeper ring of small stages wins
#ifndef SCORE_TL_KC
#define SCORE_TL_KC 16
#endif
#ifndef SCORE_TL_NST
#define SCORE_TL_NST 8
#endif
constexpr int TL_KC = SCORE_TL_KC;      // weight rows per pipeline stage
constexpr int TL_NST = SCORE_TL_NST;    // cp.async ring depth

template <int RT>
struct TileGeom {
    static constexpr int RG = RT / 4;               // 4-row groups = lanes along the rows
    static constexpr int CGW = 32 / RG;             // 4-column groups per warp
    static constexpr int NP = (TL_CT / 32) * CGW * 4;   // columns one pass covers (RT=32: 128, RT=16: 256)
    static constexpr int SLOTS = NP / 4;            // 16-byte slots per staged weight row
    static constexpr int WBUF = TL_NST * TL_KC * NP;    // floats of weight staging
};

// this thread's micro-tile: rows [4*rg, 4*rg+4) of the tile, columns [4*cg, 4*cg+4) of the pass
template <int RT>
__device__ __forceinline__ void tile_coords(int tid, int& rg, int& cg) {
    const int lane = tid & 31, warp = tid >> 5;
    rg = lane % TileGeom<RT>::RG;
    cg = warp * TileGeom<RT>::CGW + lane / TileGeom<RT>::RG;
}

__device__ __forceinline__ void tile_zero(float (&acc)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// acc[i][j] += sum_{k<K} Xs[k*RT + 4*rg + i] * W[k*ldw + 4*cg + j]      (all TL_CT threads must call)
//   Xs : shared, [K][RT]        W : global, row-major, N (<= NP, multiple of 4) columns used, ldw % 4 == 0, 16-byte aligned
//   wbuf : TileGeom<RT>::WBUF floats of shared memory
// Ends with a barrier: the caller may overwrite Xs / wbuf right away.
template <int RT>
__device__ __forceinline__ void tile_layer(const float* __restrict__ Xs, int K, const float* __restrict__ W, int ldw,
                                           int N, float (&acc)[4][4], float* wbuf, int tid) {
    using G = TileGeom<RT>;
    const int n4 = N >> 2;
    int rg, cg;
    tile_coords<RT>(tid, rg, cg);
    const bool active = cg < n4;
    auto stage = [&](int c) {
        const int k0 = c * TL_KC;
        if (k0 < K) {
            float* dst = wbuf + (c % TL_NST) * (TL_KC * G::NP);
            const int kmax = min(TL_KC, K - k0);
#pragma unroll
            for (int it = 0; it < TL_KC * G::SLOTS / TL_CT; ++it) {
                const int i = tid + it * TL_CT;
                const int k = i / G::SLOTS, c4 = i % G::SLOTS;
                if (k < kmax && c4 < n4) cp_async16(dst + k * G::NP + c4 * 4, W + (int64_t)(k0 + k) * ldw + c4 * 4, 16);
            }
        }
        cp_async_commit();
    };
    const int nchunks = (K + TL_KC - 1) / TL_KC;
#pragma unroll
    for (int c = 0; c < TL_NST - 1; ++c) stage(c);
    for (int c = 0; c < nchunks; ++c) {
        cp_async_wait<TL_NST - 2>();
        __syncthreads();
        stage(c + TL_NST - 1);   // refills the slot read in iteration c-1 (every thread is past the barrier)
        if (active) {
            const float* wb = wbuf + (c % TL_NST) * (TL_KC * G::NP) + cg * 4;
            const float* xb = Xs + (c * TL_KC) * RT + rg * 4;
            const int kmax = min(TL_KC, K - c * TL_KC);
            if (kmax == TL_KC) {
#pragma unroll
                for (int k = 0; k < TL_KC; ++k) {
                    const float4 x = *reinterpret_cast<const float4*>(xb + k * RT);
                    const float4 w = *reinterpret_cast<const float4*>(wb + k * G::NP);
                    acc[0][0] = fmaf(x.x, w.x, acc[0][0]); acc[0][1] = fmaf(x.x, w.y, acc[0][1]);
                    acc[0][2] = fmaf(x.x, w.z, acc[0][2]); acc[0][3] = fmaf(x.x, w.w, acc[0][3]);
                    acc[1][0] = fmaf(x.y, w.x, acc[1][0]); acc[1][1] = fmaf(x.y, w.y, acc[1][1]);
                    acc[1][2] = fmaf(x.y, w.z, acc[1][2]); acc[1][3] = fmaf(x.y, w.w, acc[1][3]);
                    acc[2][0] = fmaf(x.z, w.x, acc[2][0]); acc[2][1] = fmaf(x.z, w.y, acc[2][1]);
                    acc[2][2] = fmaf(x.z, w.z, acc[2][2]); acc[2][3] = fmaf(x.z, w.w, acc[2][3]);
                    acc[3][0] = fmaf(x.w, w.x, acc[3][0]); acc[3][1] = fmaf(x.w, w.y, acc[3][1]);
                    acc[3][2] = fmaf(x.w, w.z, acc[3][2]); acc[3][3] = fmaf(x.w, w.w, acc[3][3]);
                }
            } else {
                for (int k = 0; k < kmax; ++k) {
                    const float4 x = *reinterpret_cast<const float4*>(xb + k * RT);
                    const float4 w = *reinterpret_cast<const float4*>(wb + k * G::NP);
                    acc[0][0] = fmaf(x.x, w.x, acc[0][0]); acc[0][1] = fmaf(x.x, w.y, acc[0][1]);
                    acc[0][2] = fmaf(x.x, w.z, acc[0][2]); acc[0][3] = fmaf(x.x, w.w, acc[0][3]);
                    acc[1][0] = fmaf(x.y, w.x, acc[1][0]); acc[1][1] = fmaf(x.y, w.y, acc[1][1]);
                    acc[1][2] = fmaf(x.y, w.z, acc[1][2]); acc[1][3] = fmaf(x.y, w.w, acc[1][3]);
                    acc[2][0] = fmaf(x.z, w.x, acc[2][0]); acc[2][1] = fmaf(x.z, w.y, acc[2][1]);
                    acc[2][2] = fmaf(x.z, w.z, acc[2][2]); acc[2][3] = fmaf(x.z, w.w, acc[2][3]);
                    acc[3][0] = fmaf(x.w, w.x, acc[3][0]); acc[3][1] = fmaf(x.w, w.y, acc[3][1]);
                    acc[3][2] = fmaf(x.w, w.z, acc[3][2]); acc[3][3] = fmaf(x.w, w.w, acc[3][3]);
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

// store this thread's micro-tile into a k-major activation tile (the next layer's Xs): dst[(4cg+j)*RT + 4rg + i]
template <int RT>
__device__ __forceinline__ void tile_store_smem(float* dst, const float (&acc)[4][4], int rg, int cg) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
        *reinterpret_cast<float4*>(dst + (cg * 4 + j) * RT + rg * 4) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
}

// Transposing load of a [rows][width] row-major global block into a k-major tile: lane = row (conflict-free shared
// stores), each thread fetches 16 bytes of one row.  f(r, c4, v) post-processes / redirects the float4 of row r,
// columns [4*c4, 4*c4+4).  width % 4 == 0.
template <int RT, typename F>
__device__ __forceinline__ void tile_for_each_chunk(int width, int tid, F f) {
    const int w4 = width >> 2;
    for (int idx = tid; idx < RT * w4; idx += TL_CT) f(idx % RT, idx / RT);
}
__device__ __forceinline__ void tile_put4(float* Xs, int RT, int c, int r, const float4& v) {
    Xs[(c + 0) * RT + r] = v.x; Xs[(c + 1) * RT + r] = v.y; Xs[(c + 2) * RT + r] = v.z; Xs[(c + 3) * RT + r] = v.w;
}

}  // namespace score
