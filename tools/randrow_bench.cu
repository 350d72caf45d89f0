// Practical roofline of RANDOM row access on this part (standalone; nvcc -arch=sm_100a -O3 -o tools/bin/randrow_bench).
// The graded kernels of the path move 64-byte embedding rows (d = 16) / 192-byte optimizer records at random addresses;
// MEASURED_PEAKS.json's HBM figure is a streaming copy.  This program measures what random rows reach, as a function of
// row size and launch size, with the access patterns the kernels can choose from:
//   ldg     group of row_bytes/16 lanes per row, LDG.128, `depth` independent rows in flight per group
//   cpasync warp-per-tile, 16-byte cp.async.cg into shared memory, commit/wait per tile, `depth` tiles in flight
//   bulk    one cp.async.bulk (UBLKCP) per row issued per lane + mbarrier complete_tx
//   rmw     read a record, modify, write it back (the optimizer's pattern)
// Output: one line per (pattern, row bytes, rows per launch, depth): microseconds, GB/s of useful bytes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include <functional>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int LPR, int DEPTH>
__global__ void __launch_bounds__(256) ldg_kernel(const float4* __restrict__ tab, const int32_t* __restrict__ idx, int64_t n,
                                                  float4* __restrict__ out) {
    const int sub = threadIdx.x % LPR;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    const int64_t ng = (int64_t)gridDim.x * blockDim.x / LPR;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t g = g0 * DEPTH; g < n; g += ng * DEPTH) {
        int32_t id[DEPTH];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) id[u] = g + u < n ? idx[g + u] : 0;
        float4 v[DEPTH];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) v[u] = tab[(int64_t)id[u] * LPR + sub];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// read-modify-write of a record var | m | v (3 * LPR 16-byte chunks) by LPR lanes; STRIDE = chunks between records
// (3 * LPR: packed; 4 * LPR: 256-byte-aligned records for d = 16, whose 4th quarter holds the row's last_step word);
// MODE 0 read + write, 1 read only, 2 write only; EXTRA: lane 0 also reads (and writes) one int in the 4th quarter
template <int LPR, int STRIDE, int MODE, int EXTRA>
__global__ void __launch_bounds__(256) rmw_kernel(float4* __restrict__ tab, const int32_t* __restrict__ idx, int64_t n,
                                                  float4* __restrict__ out) {
    const int sub = threadIdx.x % LPR;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    const int64_t ng = (int64_t)gridDim.x * blockDim.x / LPR;
    float4 sink = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t g = g0; g < n; g += ng) {
        const int64_t base = (int64_t)idx[g] * STRIDE + sub;
        float4 a = make_float4(1.f, 2.f, 3.f, 4.f), b = a, c = a;
        int ls = 0;
        if (MODE != 2) {
            a = tab[base]; b = tab[base + LPR]; c = tab[base + 2 * LPR];
            if (EXTRA && sub == 0) ls = reinterpret_cast<const int*>(tab + base + 3 * LPR)[0];
        }
        a.x += 1.f; b.y += 1.f; c.z += a.x * 1e-9f + (float)ls;
        if (MODE != 1) {
            tab[base] = a; tab[base + LPR] = b; tab[base + 2 * LPR] = c;
            if (EXTRA && sub == 0) reinterpret_cast<int*>(tab + base + 3 * LPR)[0] = ls + 1;
        } else { sink.x += a.x + b.y + c.z; }
    }
    if (MODE == 1) out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = sink;
}
// gather of ROW-byte rows that lie STRIDE 16-byte chunks apart (the forward gather reads the var field of the record)
template <int LPR, int STRIDE>
__global__ void __launch_bounds__(256) ldg_stride_kernel(const float4* __restrict__ tab, const int32_t* __restrict__ idx, int64_t n,
                                                         float4* __restrict__ out) {
    const int sub = threadIdx.x % LPR;
    const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    const int64_t ng = (int64_t)gridDim.x * blockDim.x / LPR;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t g = g0; g < n; g += ng) {
        const float4 v = tab[(int64_t)idx[g] * STRIDE + sub];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
// warp-per-tile of TILE_ROWS rows; DEPTH tiles in flight per warp (ring in shared memory)
template <int LPR, int TILE_ROWS, int DEPTH>
__global__ void __launch_bounds__(128) cpasync_kernel(const float4* __restrict__ tab, const int32_t* __restrict__ idx, int64_t n,
                                                      float4* __restrict__ out) {
    extern __shared__ __align__(16) float4 sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    constexpr int CH = TILE_ROWS * LPR;   // chunks per tile
    float4* ring = sm + (size_t)warp * DEPTH * CH;
    const int64_t ntiles = (n + TILE_ROWS - 1) / TILE_ROWS;
    const int64_t w0 = (int64_t)blockIdx.x * warps + warp, nw = (int64_t)gridDim.x * warps;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    auto issue = [&](int64_t tile, int slot) {
        if (tile < ntiles) {
            const int32_t* ks = idx + tile * TILE_ROWS;
            for (int q = lane; q < CH; q += 32) {
                const int r = q / LPR, c = q % LPR;
                const int64_t row = tile * TILE_ROWS + r < n ? ks[r] : 0;
                cp_async16(ring + slot * CH + q, tab + row * LPR + c);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    int64_t t = w0;
#pragma unroll
    for (int s = 0; s < DEPTH - 1; ++s) issue(t + s * nw, s);
    int slot = 0;
    for (; t < ntiles; t += nw) {
        issue(t + (DEPTH - 1) * nw, (slot + DEPTH - 1) % DEPTH);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(DEPTH - 1));
        __syncwarp();
        for (int q = lane; q < CH; q += 32) { const float4 v = ring[slot * CH + q]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        __syncwarp();
        slot = (slot + 1) % DEPTH;
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// one cp.async.bulk per row, issued by the lane that owns the row; one mbarrier per warp per tile
template <int ROW_BYTES, int TILE_ROWS>
__global__ void __launch_bounds__(128) bulk_kernel(const char* __restrict__ tab, const int32_t* __restrict__ idx, int64_t n,
                                                   float4* __restrict__ out) {
    extern __shared__ __align__(128) char smc[];
    __shared__ __align__(8) uint64_t bars[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
    char* buf = smc + (size_t)warp * TILE_ROWS * ROW_BYTES;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[warp]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar));
    __syncwarp();
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    const int64_t ntiles = (n + TILE_ROWS - 1) / TILE_ROWS;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned phase = 0;
    for (int64_t t = (int64_t)blockIdx.x * warps + warp; t < ntiles; t += (int64_t)gridDim.x * warps) {
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(TILE_ROWS * ROW_BYTES) : "memory");
        __syncwarp();
        for (int r = lane; r < TILE_ROWS; r += 32) {
            const int64_t row = t * TILE_ROWS + r < n ? idx[t * TILE_ROWS + r] : 0;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(buf + r * ROW_BYTES);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                         ::"r"(dst), "l"(tab + row * ROW_BYTES), "r"(ROW_BYTES), "r"(bar) : "memory");
        }
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(bar), "r"(phase) : "memory");
        }
        phase ^= 1;
        const float4* b4 = reinterpret_cast<const float4*>(buf);
        for (int q = lane; q < TILE_ROWS * ROW_BYTES / 16; q += 32) { const float4 v = b4[q]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        __syncwarp();
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

static float time_it(cudaStream_t st, int reps, const std::function<void(int)>& f) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    if (getenv("RANDROW_REPS")) reps = atoi(getenv("RANDROW_REPS"));   // ncu pass: few launches per variant
    for (int i = 0; i < 3; ++i) f(i);
    CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(a, st));
    for (int i = 0; i < reps; ++i) f(3 + i);
    CK(cudaEventRecord(b, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms * 1000.f / reps;
}

int main(int argc, char** argv) {
    const int64_t table_bytes = (int64_t)1 << 30;   // >> 126 MB of L2
    const int NB = 16;                               // rotating index lists
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    char* tab; CK(cudaMalloc(&tab, table_bytes)); CK(cudaMemset(tab, 0, table_bytes));
    float4* out; CK(cudaMalloc(&out, (size_t)148 * 64 * 256 * sizeof(float4)));
    const int64_t sizes[] = {334232, 1 << 20, 5347712};   // Taobao B=1024 live ids, 1 M, Taobao B=16384
    const int64_t nmax = sizes[2];
    int32_t* idx; CK(cudaMalloc(&idx, sizeof(int32_t) * nmax * NB));
    std::vector<int32_t> h(nmax * NB);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("# SMs %d; table 1 GiB; %d rotating index lists; useful bytes = rows x row_bytes (+4 per id)\n", sms, NB);

    auto fill = [&](int64_t rows_in_table) {
        uint64_t s = 88172645463325252ull;
        for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (int32_t)(s % (uint64_t)rows_in_table); }
        CK(cudaMemcpy(idx, h.data(), sizeof(int32_t) * h.size(), cudaMemcpyHostToDevice));
    };
    auto report = [&](const char* pat, int row_bytes, int64_t n, int depth, float us, int rw) {
        const double bytes = (double)n * (row_bytes * rw + 4);
        printf("%-8s row %4d B  rows %8lld  depth %d  %8.2f us  %8.1f GB/s  (%.3f of 6538)\n", pat, row_bytes, (long long)n, depth, us,
               bytes / us * 1e-3, bytes / us * 1e-3 / 6538.0);
        fflush(stdout);
    };
#define RUN_LDG(LPR, DEPTH, n)                                                                                        \
    {                                                                                                                 \
        int64_t want = ((n + DEPTH - 1) / DEPTH * LPR + 255) / 256, cap = (int64_t)sms * 8;                            \
        unsigned grid = (unsigned)(want < cap ? want : cap);                                                          \
        float us = time_it(st, 40, [&](int i) { ldg_kernel<LPR, DEPTH><<<grid, 256, 0, st>>>((const float4*)tab, idx + (i % NB) * nmax, n, out); }); \
        report("ldg", LPR * 16, n, DEPTH, us, 1);                                                                     \
    }
#define RUN_CPA(LPR, TR, DEPTH, n)                                                                                    \
    {                                                                                                                 \
        size_t smem = (size_t)4 * DEPTH * TR * LPR * 16;                                                              \
        CK(cudaFuncSetAttribute(cpasync_kernel<LPR, TR, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cpasync_kernel<LPR, TR, DEPTH>, 128, smem)); \
        int64_t want = ((n + TR - 1) / TR + 3) / 4, cap = (int64_t)sms * occ;                                          \
        unsigned grid = (unsigned)(want < cap ? want : cap);                                                          \
        float us = time_it(st, 40, [&](int i) { cpasync_kernel<LPR, TR, DEPTH><<<grid, 128, smem, st>>>((const float4*)tab, idx + (i % NB) * nmax, n, out); }); \
        printf("  [tile %d rows, %d CTAs/SM] ", TR, occ); report("cpasync", LPR * 16, n, DEPTH, us, 1);               \
    }
#define RUN_BULK(RB, TR, n)                                                                                           \
    {                                                                                                                 \
        size_t smem = (size_t)4 * TR * RB;                                                                            \
        CK(cudaFuncSetAttribute(bulk_kernel<RB, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bulk_kernel<RB, TR>, 128, smem));         \
        int64_t want = ((n + TR - 1) / TR + 3) / 4, cap = (int64_t)sms * occ;                                          \
        unsigned grid = (unsigned)(want < cap ? want : cap);                                                          \
        float us = time_it(st, 40, [&](int i) { bulk_kernel<RB, TR><<<grid, 128, smem, st>>>(tab, idx + (i % NB) * nmax, n, out); }); \
        printf("  [tile %d rows, %d CTAs/SM] ", TR, occ); report("bulk", RB, n, 1, us, 1);                             \
    }
#define RUN_RMW(LPR, STRIDE, MODE, EXTRA, n)                                                                          \
    {                                                                                                                 \
        int64_t want = (n * LPR + 255) / 256, cap = (int64_t)sms * 8;                                                  \
        unsigned grid = (unsigned)(want < cap ? want : cap);                                                          \
        float us = time_it(st, 40, [&](int i) { rmw_kernel<LPR, STRIDE, MODE, EXTRA><<<grid, 256, 0, st>>>((float4*)tab, idx + (i % NB) * nmax, n, out); }); \
        printf("  [record stride %d B%s] ", STRIDE * 16, EXTRA ? " + last_step word" : "");                            \
        report(MODE == 0 ? "rmw" : MODE == 1 ? "rec-read" : "rec-write", LPR * 48, n, 1, us, MODE == 0 ? 2 : 1);       \
    }
#define RUN_LDGS(LPR, STRIDE, n)                                                                                      \
    {                                                                                                                 \
        int64_t want = (n * LPR + 255) / 256, cap = (int64_t)sms * 8;                                                  \
        unsigned grid = (unsigned)(want < cap ? want : cap);                                                          \
        float us = time_it(st, 40, [&](int i) { ldg_stride_kernel<LPR, STRIDE><<<grid, 256, 0, st>>>((const float4*)tab, idx + (i % NB) * nmax, n, out); }); \
        printf("  [row stride %d B] ", STRIDE * 16); report("ldg", LPR * 16, n, 1, us, 1);                             \
    }
    // ---- 64-byte rows at stride 192 bytes is what the forward gather sees (var field of the record); plain 64-byte
    //      stride is the densest case.  Both cost a 128-byte DRAM access per row on this part (profiles/README.md).
    fill(table_bytes / 64);
    for (int64_t n : sizes) {
        RUN_LDG(4, 1, n); RUN_LDG(4, 2, n); RUN_LDG(4, 4, n); RUN_LDG(4, 8, n);
        RUN_CPA(4, 60, 1, n); RUN_CPA(4, 60, 2, n); RUN_CPA(4, 60, 3, n);
        RUN_BULK(64, 60, n);
    }
    fill(table_bytes / 256);
    for (int64_t n : sizes) {
        const int64_t m = n / 4;
        RUN_LDG(16, 1, m); RUN_LDG(16, 2, m); RUN_LDG(16, 4, m);
        RUN_CPA(16, 40, 1, m); RUN_CPA(16, 40, 2, m);
        RUN_BULK(256, 40, m);
    }
    // ---- 64-byte rows inside records of 192 B (packed var|m|v) and 256 B (aligned)
    fill(table_bytes / 256);
    for (int64_t n : sizes) { RUN_LDGS(4, 12, n); RUN_LDGS(4, 16, n); }
    // ---- optimizer records (d = 16): packed 192 B vs 256-byte-aligned, read + write / read only / write only
    for (int64_t n : {(int64_t)128337, (int64_t)1 << 20}) {
        RUN_RMW(4, 12, 0, 0, n); RUN_RMW(4, 16, 0, 0, n); RUN_RMW(4, 16, 0, 1, n);
        RUN_RMW(4, 12, 1, 0, n); RUN_RMW(4, 16, 1, 0, n);
        RUN_RMW(4, 12, 2, 0, n); RUN_RMW(4, 16, 2, 0, n);
    }
    // ---- d = 64: 768-byte records
    fill(table_bytes / 768);
    for (int64_t n : {(int64_t)128337 / 4, (int64_t)1 << 18}) RUN_RMW(16, 48, 0, 0, n);
    return 0;
}
