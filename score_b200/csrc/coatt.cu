// Lean instances of the fused gather + co-attention kernels for the compiled-in geometries of the reference's data
// sets (class SCORE and the two single-side ablations: both co-attentions active, K <= 16, d a multiple of 16).
//
// Same arithmetic and position layout as embed.cu (see its header: rank-1 form of co_attention, score.py:147-167;
// slice-major positions); what changes is the instruction budget.  ncu on the general kernel (profiles/r02g): 830 warp
// instructions per live Taobao slice for 3.8 KB of rows - a third of them 64-bit address arithmetic, predicated shuffle
// trees and per-chunk index math - against a random-row DRAM ceiling of 12.4 us per launch (tools/randrow_bench.cu).
// Here
//   * the warps walk the list of LIVE slices (t < length, built by build_keys): a dead slice costs nothing and the
//     Taobao batch fits one resident wave (8 CTAs x 4 warps per SM: 64 registers);
//   * every shared-memory access is lane base + immediate (the chunk -> (segment, neighbor, field) maps are compile
//     time), the co-attention kernel chunk of a lane sits in registers whenever 32 is a multiple of the chunks per
//     neighbor;
//   * the dot products stay per-chunk partials in shared memory and are summed by the lane that owns the neighbor
//     (no shuffle tree per row);
//   * backward keeps the warp's share of the co-attention kernel gradient in registers across its slices.
// Any other configuration (run-time geometry, K > 16, RCA / RRN sum pooling) takes the general kernels of embed.cu.
// Measured on B200 (profiles/r2n_*): 5.19 M -> 3.33 M warp instructions per launch, K-A alone 15.3-16.5 -> 12.4-13.7 us
// (a kernel that only gathers the same rows: 12.4 us), K-D 17.7 -> 14.3-15.7 us; SCORE_COATT_LEAN=0 selects the general kernels.
#include "kernels.h"

namespace score {

namespace {

template <int K_, int FI_, int FU_, int D_>
struct LG {
    static constexpr int K = K_, FI = FI_, FU = FU_, D = D_;
    static constexpr int CPR = D / 4;                       // 16-byte chunks per row
    static constexpr int NFI = K * FI, NFU = K * FU, NROWS = 2 * NFI + 2 * NFU;
    static constexpr int NCH = NROWS * CPR;                 // chunks per slice
    static constexpr int NIT = (NCH + 31) / 32;
    static constexpr int DI = FI * D, DU = FU * D, DS = DI + DU;
    static constexpr int NW = 3 * DI + 3 * DU;              // floats of both co-attention kernels
    // per-warp shared memory (floats): rows | part | wts
    static constexpr int ROWS_F = NCH * 4, PART_F = NCH, WTS_F = 64;
    static constexpr int WARP_F = ROWS_F + PART_F + WTS_F;
    // segment s: first row, rows, fields per node, chunk offset, chunks per neighbor, float offset of its kernel slice
    __host__ __device__ static constexpr int row0(int s) { return s == 0 ? 0 : s == 1 ? NFI : s == 2 ? 2 * NFI : 2 * NFI + NFU; }
    __host__ __device__ static constexpr int F(int s) { return s < 2 ? FI : FU; }
    __host__ __device__ static constexpr int ch0(int s) { return row0(s) * CPR; }
    __host__ __device__ static constexpr int fm(int s) { return F(s) * CPR; }
    __host__ __device__ static constexpr int nch(int s) { return K * F(s) * CPR; }
    __host__ __device__ static constexpr int wofs(int s) { return s == 0 ? DI : s == 1 ? 2 * DI : s == 2 ? 3 * DI + DU : 3 * DI + 2 * DU; }
};

__device__ __forceinline__ float dot4f(const float4& a, const float4& b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float hsum16(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float hmax16(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

// all chunks of one slice: lane q mod 32 of iteration q / 32 fetches chunk q (row q / CPR, 16 bytes at 4 * (q mod CPR))
template <class G>
__device__ __forceinline__ void lean_gather(float* rows, const float* __restrict__ emb, int es, const int32_t* __restrict__ ks, int lane) {
    const int32_t* kl = ks + lane / G::CPR;
    const float* el = emb + (lane % G::CPR) * 4;
    float* dl = rows + lane * 4;
#pragma unroll
    for (int j = 0; j < G::NIT; ++j) {
        if (j * 32 + 31 < G::NCH || j * 32 + lane < G::NCH) {
            const int32_t id = __ldg(kl + j * (32 / G::CPR));
            cp_async16(dl + j * 128, el + (int64_t)id * es, id != 0 ? 16 : 0);   // id 0: zero-fill, row 0 is never read
        }
    }
    cp_async_commit();
}

// part[q] = <chunk q of the slice, matching chunk of vec(segment)>; vec = co-attention kernel slice (forward) or the
// gradient of the segment's pooled output (backward).  VREG: the lane's chunk of vec is the same in every iteration.
template <class G, int S>
__device__ __forceinline__ void lean_seg_dots(const float* rows, const float* vec, float* part, int lane) {
    constexpr int FM = G::fm(S), NC = G::nch(S), C0 = G::ch0(S);
    const float4* r4 = reinterpret_cast<const float4*>(rows) + C0 + lane;
    float* pl = part + C0 + lane;
    if (32 % FM == 0) {
        const float4 v = *reinterpret_cast<const float4*>(vec + (lane % FM) * 4);
#pragma unroll
        for (int q0 = 0; q0 < NC; q0 += 32)
            if (q0 + 31 < NC || q0 + lane < NC) pl[q0] = dot4f(r4[q0], v);
    } else {
        int m = lane % FM;
#pragma unroll
        for (int q0 = 0; q0 < NC; q0 += 32) {
            if (q0 + 31 < NC || q0 + lane < NC) pl[q0] = dot4f(r4[q0], *reinterpret_cast<const float4*>(vec + m * 4));
            m += 32 % FM;
            if (m >= FM) m -= FM;
        }
    }
}

// sum of the CNT partials part[0..CNT) (CNT a multiple of 4, 16-byte aligned)
template <int CNT>
__device__ __forceinline__ float lean_sum_part(const float* part) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CNT; c += 4) {
        const float4 v = *reinterpret_cast<const float4*>(part + c);
        s += (v.x + v.y) + (v.z + v.w);
    }
    return s;
}

template <class G>
__global__ void __launch_bounds__(128, 8) coatt_fwd_lean_kernel(Dims dm, CoattArgs a, const int32_t* __restrict__ live, int M) {
    pdl_enter();
    extern __shared__ __align__(16) float sm[];
    constexpr int K = G::K, CPR = G::CPR, DI = G::DI, DU = G::DU, DS = G::DS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Wsm = sm;
    for (int i = threadIdx.x; i < 3 * DI; i += blockDim.x) Wsm[i] = a.w_item[i];
    for (int i = threadIdx.x; i < 3 * DU; i += blockDim.x) Wsm[3 * DI + i] = a.w_user[i];
    float* rows = sm + ((G::NW + 3) & ~3) + warp * G::WARP_F;
    float* part = rows + G::ROWS_F;
    float* wts = part + G::PART_F;       // w1[16] | w2[16] | 1/K [16]: per-neighbor weights of the pooled outputs
    const float invK = 1.0f / (float)K;
    if (lane < 16) { wts[lane] = 0.f; wts[16 + lane] = 0.f; wts[32 + lane] = invK; }
    __syncthreads();
    const int nlive = live[M];
    const int half = lane >> 4, i = lane & 15;
    const bool act = i < K;
    // pooling: lane e4 < 2*DS/4 owns one 16-byte output chunk
    //   user_side = [sum_i w1_i user_1hop[i] | sum_i w2_i user_2hop[i]],  item_side = [mean item_1hop | mean item_2hop]
    constexpr int NOUT = (2 * DS) / 4, NOIT = (NOUT + 31) / 32;
    int p_src[NOIT], p_stride[NOIT], p_w[NOIT], p_dst[NOIT];   // chunk index of neighbor 0, chunks per neighbor, weights, output
#pragma unroll
    for (int u = 0; u < NOIT; ++u) {
        const int e = (u * 32 + lane) * 4;
        int s, c, wsel, dst;
        if (e < DI) { s = 0; c = e; wsel = 0; dst = e; }                                   // user side, 1-hop part
        else if (e < DS) { s = 2; c = e - DI; wsel = 16; dst = e; }                        // user side, 2-hop part
        else if (e < DS + DU) { s = 3; c = e - DS; wsel = 32; dst = DS + c; }              // item side, 1-hop part
        else { s = 1; c = e - DS - DU; wsel = 32; dst = DS + DU + c; }                     // item side, 2-hop part
        p_src[u] = (s == 0 ? G::ch0(0) : s == 1 ? G::ch0(1) : s == 2 ? G::ch0(2) : G::ch0(3)) + (c >> 2);
        p_stride[u] = s < 2 ? G::fm(0) : G::fm(2);
        p_w[u] = wsel; p_dst[u] = dst;
    }
    const int wstride = gridDim.x * 4;
    for (int idx = blockIdx.x * 4 + warp; idx < nlive; idx += wstride) {
        const int ent = live[idx];                   // (b << 8) | t
        const int b = ent >> 8, slice = b * dm.T + (ent & 255);
        lean_gather<G>(rows, a.emb, a.es, a.keys + (int64_t)slice * G::NROWS, lane);
        const float cz = half ? a.c_user[b] : a.c_item[b];
        float* xu_g = a.xhg_u + (int64_t)slice * dm.ldxs[0]; float* xu_c = a.xhc_u + (int64_t)slice * dm.ldxs[0];
        float* xi_g = a.xhg_i + (int64_t)slice * dm.ldxs[1]; float* xi_c = a.xhc_i + (int64_t)slice * dm.ldxs[1];
        float* info = a.key + (int64_t)slice * a.ldkey + a.key_off + half * 2 * K;
        float* sr = a.save_r + (int64_t)slice * 2 * K + half * K; float* sw = a.save_w + (int64_t)slice * 2 * K + half * K;
        cp_async_wait<0>();
        __syncwarp();
        // (1) per-chunk partial dot products with the co-attention kernels (W1 for seq1 rows, W2 for seq2 rows)
        lean_seg_dots<G, 0>(rows, Wsm + G::wofs(0), part, lane);
        lean_seg_dots<G, 1>(rows, Wsm + G::wofs(1), part, lane);
        lean_seg_dots<G, 2>(rows, Wsm + G::wofs(2), part, lane);
        lean_seg_dots<G, 3>(rows, Wsm + G::wofs(3), part, lane);
        __syncwarp();
        // (2) relatedness r_i = relu(target part + seq1[i] part + seq2[i] part); softmax over the K neighbors.
        //     lanes 0..15: co-attention #1 (user_1hop, item_2hop, target_item); lanes 16..31: #2 (user_2hop, item_1hop, target_user)
        float z = cz;
        if (act) {
            if (half == 0) z += lean_sum_part<G::fm(0)>(part + G::ch0(0) + i * G::fm(0)) + lean_sum_part<G::fm(1)>(part + G::ch0(1) + i * G::fm(1));
            else z += lean_sum_part<G::fm(2)>(part + G::ch0(2) + i * G::fm(2)) + lean_sum_part<G::fm(3)>(part + G::ch0(3) + i * G::fm(3));
        }
        const float r = act ? fmaxf(z, 0.f) : 0.f;
        const float mx = hmax16(r);                 // r >= 0: the padding lanes (0) never exceed the maximum
        const float e = act ? expf(r - mx) : 0.f;
        const float w = e / hsum16(e);
        const float s = hsum16(r);
        if (act) {
            wts[half * 16 + i] = w;
            info[i] = (float)K * r;                 // atten_info (score.py:165-166)
            info[K + i] = s;
            sr[i] = r; sw[i] = w;
        }
        __syncwarp();
        // (3) pooling
        const float4* r4 = reinterpret_cast<const float4*>(rows);
#pragma unroll
        for (int u = 0; u < NOIT; ++u) {
            if (u * 32 + 31 < NOUT || u * 32 + lane < NOUT) {
                const float4* src = r4 + p_src[u];
                const float* wv = wts + p_w[u];
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int n = 0; n < K; ++n) {
                    const float4 v = src[n * p_stride[u]];
                    const float wn = wv[n];
                    acc.x = fmaf(wn, v.x, acc.x); acc.y = fmaf(wn, v.y, acc.y); acc.z = fmaf(wn, v.z, acc.z); acc.w = fmaf(wn, v.w, acc.w);
                }
                const int dst = p_dst[u];
                float* d0 = dst < DS ? xu_g + dst : xi_g + (dst - DS);
                float* d1 = dst < DS ? xu_c + dst : xi_c + (dst - DS);
                *reinterpret_cast<float4*>(d0) = acc;
                *reinterpret_cast<float4*>(d1) = acc;
            }
        }
        __syncwarp();
    }
    (void)CPR;
}

// backward of both co-attentions of one live slice (formulas: embed.cu, coatt_bwd_kernel)
template <class G>
__global__ void __launch_bounds__(128, 6) coatt_bwd_lean_kernel(Dims dm, CoattBwdArgs a, const int32_t* __restrict__ live, int M) {
    pdl_enter();
    extern __shared__ __align__(16) float sm[];
    constexpr int K = G::K, DI = G::DI, DU = G::DU, DS = G::DS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* Wsm = sm;
    for (int i = threadIdx.x; i < 3 * DI; i += blockDim.x) Wsm[i] = a.w_item[i];
    for (int i = threadIdx.x; i < 3 * DU; i += blockDim.x) Wsm[3 * DI + i] = a.w_user[i];
    constexpr int DBUF_F = 2 * DS;                              // d user_side [DS] | d item_side [DS]
    constexpr int NACC = 2 * DI + 2 * DU, NACC4 = NACC / 4, NAIT = (NACC4 + 31) / 32;
    constexpr int BW_F = G::WARP_F + DBUF_F;
    float* rows = sm + ((G::NW + 3) & ~3) + warp * BW_F;
    float* part = rows + G::ROWS_F;
    float* wts = part + G::PART_F;       // w1[16] | w2[16] | dz1[16] | dz2[16]
    float* dbuf = wts + G::WTS_F;
    wts[lane] = 0.f; wts[32 + lane] = 0.f;
    __syncthreads();
    const int nlive = live[M];
    const int half = lane >> 4, i = lane & 15;
    const bool act = i < K;
    const float invK = 1.0f / (float)K;
    // co-attention kernel gradient: lane e4 < NACC/4 owns one 16-byte chunk of [dW1_item | dW2_item | dW1_user | dW2_user]
    float4 acc[NAIT];
    int a_src[NAIT], a_stride[NAIT], a_z[NAIT];
#pragma unroll
    for (int u = 0; u < NAIT; ++u) {
        acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        const int e = (u * 32 + lane) * 4;
        int s, c;
        if (e < DI) { s = 0; c = e; } else if (e < 2 * DI) { s = 1; c = e - DI; } else if (e < 2 * DI + DU) { s = 2; c = e - 2 * DI; } else { s = 3; c = e - 2 * DI - DU; }
        a_src[u] = (s == 0 ? G::ch0(0) : s == 1 ? G::ch0(1) : s == 2 ? G::ch0(2) : G::ch0(3)) + (c >> 2);
        a_stride[u] = s < 2 ? G::fm(0) : G::fm(2);
        a_z[u] = s < 2 ? 32 : 48;
    }
    const int wstride = gridDim.x * 4;
    for (int idx = blockIdx.x * 4 + warp; idx < nlive; idx += wstride) {
        const int ent = live[idx];
        const int slice = (ent >> 8) * dm.T + (ent & 255);
        lean_gather<G>(rows, a.emb, a.es, a.keys + (int64_t)slice * G::NROWS, lane);
        {   // incoming gradients of this slice (overlaps the row gather)
            const float4* dxu = reinterpret_cast<const float4*>(a.dxu + (int64_t)slice * DS);
            const float4* dxi = reinterpret_cast<const float4*>(a.dxi + (int64_t)slice * DS);
            float4* d4 = reinterpret_cast<float4*>(dbuf);
#pragma unroll
            for (int c = 0; c < DS / 4; c += 32)
                if (c + 31 < DS / 4 || c + lane < DS / 4) { d4[c + lane] = dxu[c + lane]; d4[DS / 4 + c + lane] = dxi[c + lane]; }
        }
        float w_h = 0.f, r_h = 0.f, di_n = 0.f, di_t = 0.f;
        if (act) {
            w_h = a.save_w[(int64_t)slice * 2 * K + half * K + i];
            r_h = a.save_r[(int64_t)slice * 2 * K + half * K + i];
            if (a.dkey) {
                const float* dinfo = a.dkey + (int64_t)slice * a.ldkey + a.key_off + half * 2 * K;
                di_n = dinfo[i]; di_t = dinfo[K + i];
            }
        }
        cp_async_wait<0>();
        __syncwarp();
        // (1) dw_i = <dout1, seq1[i]>: per-chunk partials of the seq1 segments with their slice of d user_side
        lean_seg_dots<G, 0>(rows, dbuf, part, lane);
        lean_seg_dots<G, 2>(rows, dbuf + DI, part, lane);
        __syncwarp();
        // (2) per-neighbor scalars
        float dw = 0.f;
        if (act) dw = half == 0 ? lean_sum_part<G::fm(0)>(part + G::ch0(0) + i * G::fm(0)) : lean_sum_part<G::fm(2)>(part + G::ch0(2) + i * G::fm(2));
        const float dotw = hsum16(w_h * dw);
        const float tail = hsum16(di_t);
        float dz = 0.f;
        if (act) {
            const float dr = (float)K * di_n + tail + w_h * (dw - dotw);
            dz = r_h > 0.f ? dr : 0.f;
            wts[half * 16 + i] = w_h; wts[32 + half * 16 + i] = dz;
        }
        const float sdz = hsum16(dz);
        if (i == 0) a.sdz[(int64_t)slice * 2 + half] = sdz;
        __syncwarp();
        // (3) gradient rows of the slice's positions, one contiguous block: dseq1[i] = w_i dout1 + dz_i W1, dseq2[i] = dout2 / K + dz_i W2
        {
            float4* gout = reinterpret_cast<float4*>(a.grad_rows) + (int64_t)slice * G::NCH;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int FM = s < 2 ? G::fm(0) : G::fm(2), NC = s < 2 ? G::nch(0) : G::nch(2);
                const int C0 = s == 0 ? G::ch0(0) : s == 1 ? G::ch0(1) : s == 2 ? G::ch0(2) : G::ch0(3);
                const float* dv = dbuf + (s == 0 ? 0 : s == 1 ? DS + DU : s == 2 ? DI : DS);
                const float* wv = Wsm + (s == 0 ? G::wofs(0) : s == 1 ? G::wofs(1) : s == 2 ? G::wofs(2) : G::wofs(3));
                const float* ai = wts + (s == 0 ? 0 : 16);
                const float* dzi = wts + (s < 2 ? 32 : 48);
                int m = lane % FM, n = lane / FM;                 // chunk inside the neighbor, neighbor
#pragma unroll
                for (int q0 = 0; q0 < NC; q0 += 32) {
                    if (q0 + 31 < NC || q0 + lane < NC) {
                        const float a_i = (s == 0 || s == 2) ? ai[n] : invK;
                        const float dzn = dzi[n];
                        const float4 dvv = *reinterpret_cast<const float4*>(dv + m * 4), wvv = *reinterpret_cast<const float4*>(wv + m * 4);
                        float4 o;
                        o.x = a_i * dvv.x + dzn * wvv.x; o.y = a_i * dvv.y + dzn * wvv.y;
                        o.z = a_i * dvv.z + dzn * wvv.z; o.w = a_i * dvv.w + dzn * wvv.w;
                        gout[C0 + q0 + lane] = o;
                    }
                    m += 32 % FM; n += 32 / FM;
                    if (m >= FM) { m -= FM; ++n; }
                }
            }
        }
        // (4) dW1 += sum_i dz_i seq1[i], dW2 += sum_i dz_i seq2[i] (registers, across the warp's slices)
        {
            const float4* r4 = reinterpret_cast<const float4*>(rows);
#pragma unroll
            for (int u = 0; u < NAIT; ++u) {
                if (u * 32 + 31 < NACC4 || u * 32 + lane < NACC4) {
                    const float4* src = r4 + a_src[u];
                    const float* zv = wts + a_z[u];
                    float4 sacc = acc[u];
#pragma unroll
                    for (int n = 0; n < K; ++n) {
                        const float4 v = src[n * a_stride[u]];
                        const float zn = zv[n];
                        sacc.x = fmaf(zn, v.x, sacc.x); sacc.y = fmaf(zn, v.y, sacc.y); sacc.z = fmaf(zn, v.z, sacc.z); sacc.w = fmaf(zn, v.w, sacc.w);
                    }
                    acc[u] = sacc;
                }
            }
        }
        __syncwarp();
    }
    // fixed-order sum over the CTA's warps -> one partial row per CTA
    __syncthreads();
    float* red = sm + ((G::NW + 3) & ~3);     // the warps are done with their regions
#pragma unroll
    for (int u = 0; u < NAIT; ++u)
        if (u * 32 + lane < NACC4) *reinterpret_cast<float4*>(red + warp * NACC + (u * 32 + lane) * 4) = acc[u];
    __syncthreads();
    for (int c = threadIdx.x; c < NACC; c += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) s += red[w * NACC + c];
        a.partials[(int64_t)blockIdx.x * NACC + c] = s;
    }
}

int lean_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <class G>
bool fwd_lean(cudaStream_t st, const Dims& dm, const CoattArgs& a, const int32_t* live) {
    const size_t smem = (size_t)(((G::NW + 3) & ~3) + 4 * G::WARP_F) * sizeof(float);
    static int per_sm = -1;
    if (per_sm < 0) {
        if (cudaFuncSetAttribute(coatt_fwd_lean_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coatt_fwd_lean_kernel<G>, 128, smem) != cudaSuccess || per_sm < 1) {
            per_sm = 0;
            cudaGetLastError();
        }
    }
    if (per_sm == 0) return false;
    const int M = dm.B * dm.T;
    int64_t want = ((int64_t)M + 3) / 4, cap = (int64_t)lean_sms() * per_sm;
    launch_chain(coatt_fwd_lean_kernel<G>, dim3((unsigned)(want < cap ? want : cap)), dim3(128), smem, st, dm, a, live, M);
    return true;
}
template <class G>
bool bwd_lean(cudaStream_t st, const Dims& dm, const CoattBwdArgs& a, const int32_t* live) {
    constexpr int NACC = 2 * G::DI + 2 * G::DU;
    constexpr int BW_F = G::WARP_F + 2 * G::DS;
    static_assert(4 * NACC <= 4 * BW_F, "the warps' regions hold the CTA reduction");
    const size_t smem = (size_t)(((G::NW + 3) & ~3) + 4 * BW_F) * sizeof(float);
    static int ok = -1;
    if (ok < 0) {
        ok = cudaFuncSetAttribute(coatt_bwd_lean_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess ? 1 : 0;
        if (!ok) cudaGetLastError();
    }
    if (!ok) return false;
    const int M = dm.B * dm.T;
    // the number of CTAs is the number of partial rows the caller reduces: every one of them is written
    launch_chain(coatt_bwd_lean_kernel<G>, dim3((unsigned)a.n_partials), dim3(128), smem, st, dm, a, live, M);
    return true;
}

bool lean_disabled() {
    static int f = -1;
    if (f < 0) { const char* e = getenv("SCORE_COATT_LEAN"); f = (e && atoi(e) == 0) ? 1 : 0; }
    return f != 0;
}

}  // namespace

#define LEAN_DISPATCH(FN, ...)                                                                        \
    do {                                                                                              \
        if (dm.K == 10 && dm.fi == 2 && dm.fu == 1 && dm.d == 16) return FN<LG<10, 2, 1, 16>>(__VA_ARGS__);   \
        if (dm.K == 10 && dm.fi == 4 && dm.fu == 3 && dm.d == 16) return FN<LG<10, 4, 3, 16>>(__VA_ARGS__);   \
        if (dm.K == 10 && dm.fi == 5 && dm.fu == 1 && dm.d == 16) return FN<LG<10, 5, 1, 16>>(__VA_ARGS__);   \
        if (dm.K == 10 && dm.fi == 1 && dm.fu == 1 && dm.d == 64) return FN<LG<10, 1, 1, 64>>(__VA_ARGS__);   \
        return false;                                                                                 \
    } while (0)

// true: the lean instance was launched; false: the caller launches the general kernel
bool try_coatt_fwd_lean(cudaStream_t st, const Dims& dm, const CoattArgs& a, const int32_t* live) {
    if (!live || a.sum_pool || dm.hop1_only || dm.T > 255 || lean_disabled()) return false;
    LEAN_DISPATCH(fwd_lean, st, dm, a, live);
}
bool try_coatt_bwd_lean(cudaStream_t st, const Dims& dm, const CoattBwdArgs& a, const int32_t* live) {
    if (!live || a.sum_pool || dm.hop1_only || dm.T > 255 || lean_disabled()) return false;
    LEAN_DISPATCH(bwd_lean, st, dm, a, live);
}

}  // namespace score
