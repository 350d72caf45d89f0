// Launcher declarations shared by the kernel translation units and model.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

namespace score {

// Every launcher bumps this (bench.py reports it as gpu_launches).
extern int64_t g_launch_count;

// Launch of a kernel of the step's critical chain, optionally with programmatic stream serialization (common.cuh:
// pdl_enter).  Measured on B200 (profiles/README.md): with every chain kernel releasing its successor at entry the
// Taobao step got 5 % SLOWER (0.430 vs 0.410 ms) - the early CTAs of the successor hold SM slots the side streams
// would have used - so plain stream order is the default and SCORE_PDL=1 turns the experiment on.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---------------------------------------------------------------- dense layers (gemm.cu)
enum GemmEpi {
    EPI_STORE = 0,        // C = acc
    EPI_BIAS = 1,         // C = acc + bias[n]
    EPI_BIAS_RELU = 2,    // C = relu(acc + bias[n])
    EPI_BIAS_RELU_DROP = 3,  // tf.nn.dropout(relu(acc + bias)) : x / keep_prob * mask   (score.py:70-73)
    EPI_MASK = 4,         // C = aux[m,n] > 0 ? acc (/ keep_prob if mask_dropout) : 0     (relu/dropout backward)
    EPI_ACCUM = 5,        // C += acc
    EPI_SPLIT = 6         // split-K partials: C + z * c_split_stride (deterministic two-stage reduce)
};

struct GemmArgs {
    const float* A; int64_t a_rs, a_cs;   // A(m,k) = A[m*a_rs + k*a_cs]
    const float* B; int64_t b_rs, b_cs;   // B(k,n) = B[k*b_rs + n*b_cs]
    float* C; int64_t c_rs;               // C(m,n) = C[m*c_rs + n]
    int M, N, K;
    int epi;
    const float* bias;
    const float* aux; int64_t aux_rs;
    int mask_dropout;
    int splits; int64_t c_split_stride;
    float* colsum; int64_t colsum_split_stride;   // EPI_SPLIT: also sum_k B(k,n) -> bias gradient
    const Hyper* hp; uint32_t rng_stream;
};
struct GemmBatch { GemmArgs g[8]; };
void launch_gemm(cudaStream_t st, const GemmArgs& a);
// up to 8 independent problems (same operand orientation, same split count) in one launch
void launch_gemm_batch(cudaStream_t st, const GemmArgs* list, int n);

// ---------------------------------------------------------------- embedding front end (embed.cu)
struct Dims {
    int64_t V;
    int B, T, K, d, H, fu, fi;
    int Du, Di, Ds, Dk, Dfc;      // user/item node width, GRU input width, key width, fc input width
    int ldx;                       // Ds + H: leading dim of the [x || h] buffers
    int Dx[2], ldxs[2];            // GRU input width / leading dim of [x || h] per side (user, item): Ds / ldx, except RRN
                                   // (slice_model.py:155-173: user side = sum of user_1hop [Di], item side = sum of item_1hop [Du])
    int hop1_only;                 // RRN: the 2-hop tensors are fed but never consumed - their positions carry key 0
    int nrows;                     // K * (2*fi + 2*fu): positions (ids) per (b,t) slice
    int64_t off_tu, off_ti, N;     // slice-major position space (embed.cu): B*T*nrows history positions, then the targets
    int model_type;
};

// Where the eight tensors of the current batch lie (device memory): the handle's staging buffers for a host batch,
// the caller's own tensors for a device batch.  Lives in device memory so a captured graph follows it.
struct BatchPtrs {
    const int32_t* u1; const int32_t* u2; const int32_t* i1; const int32_t* i2;   // user_1hop, user_2hop, item_1hop, item_2hop
    const int32_t* tu; const int32_t* ti; const int32_t* label; const int32_t* length;
};

// LAZY optimizer mode: claim the stale rows among the keys while they are built (see scatter.cu: emb_replay_kernel)
// counter: two int32; pingpong != 0: this launch appends through counter[hp->seq & 1] and zeroes the other one (no
// memset node on the critical path); pingpong == 0: counter[0], zeroed by the launcher
struct ClaimArgs { int32_t* last_step; const Hyper* hp; int32_t* list; int32_t* counter; int pingpong; };
// live (may be null): B*T + 1 ints - the live slices (t < length[b]) of the batch as (b << 8) | t, in (b, t) order, and
// their number at index B*T: the work list of the lean co-attention kernels (coatt.cu); needs T <= 255
void launch_build_keys(cudaStream_t st, const Dims& dm, const BatchPtrs* bp_dev, int32_t* keys, int32_t* label_out,
                       int32_t* length_out, int32_t* err_flag, const ClaimArgs* claim = nullptr, int32_t* live = nullptr);
// replay the rows of a claim list (2 int32 per entry: row, last step); *counter entries
void launch_emb_replay(cudaStream_t st, const int32_t* claim_list, const int32_t* claim_counter, int64_t max_rows, float* emb,
                       float* m, float* v, int d, int es, const float* alpha_hist, const Hyper* hp, int pingpong = 0);

// `es` in the argument structs below: floats between consecutive rows of the table `emb` points into - 3*d for the
// handle's own table (one record var | m | v per row, model.cu), d for a staged mini-table of a row-sharded step
struct TargetArgs {
    const float* emb; int es; const int32_t* keys;
    const float* w_item; const float* b_item;   // co-attention #1 kernel [3*Di] (rows: target|seq1|seq2), bias
    const float* w_user; const float* b_user;   // co-attention #2 kernel [3*Du]
    float* q0;      // [B, Ds]  = [target_user || target_item]      (score.py:210)
    float* fc_in; int fc_off;   // fc_in[b, fc_off:] = [target_item || target_user] (score.py:217)
    float* c_item; float* c_user;   // [B] target part of the relatedness pre-activation (+bias)
};
void launch_target_fwd(cudaStream_t st, const Dims& dm, const TargetArgs& a);

struct CoattArgs {
    const float* emb; int es; const int32_t* keys; const int32_t* length;
    const float* w_item; const float* w_user;
    const float* c_item; const float* c_user;
    float* xhg_u; float* xhc_u; float* xhg_i; float* xhc_i;   // [M, ldxs[side]], x part written here
    float* key; int ldkey; int key_off;                       // atten_info -> key[:, key_off : key_off+4K]
    float* save_r; float* save_w;                             // [M, 2K]
    int sum_pool;   // RCA (score.py:266-269): plain sum over the K neighbors, no relatedness, no atten_info
    const int32_t* live;   // live-slice list of launch_build_keys (may be null: general kernel)
};
void launch_coatt_fwd(cudaStream_t st, const Dims& dm, const CoattArgs& a);

struct CoattBwdArgs {
    const float* emb; int es; const int32_t* keys; const int32_t* length;
    const float* w_item; const float* w_user;
    const float* save_r; const float* save_w;
    const float* dxu; const float* dxi;        // [M, Dx[0]], [M, Dx[1]]
    const float* dkey; int ldkey; int key_off; // d atten_info (NULL: atten_info has no consumer - RIA)
    int sum_pool;                              // RCA: backward of the plain neighbor sum
    float* grad_rows;                          // [N, d] per-position embedding gradient rows
    float* sdz;                                // [M, 2]  sum_i d z_i per slice and co-attention
    float* partials; int n_partials;           // [n_partials, 2*Di + 2*Du] per-CTA dW1|dW2 (item), dW1|dW2 (user)
    const int32_t* live;                       // live-slice list of launch_build_keys (may be null: general kernel)
};
// lean instances for the compiled-in geometries (coatt.cu); false: not applicable, launch the general kernel
bool try_coatt_fwd_lean(cudaStream_t st, const Dims& dm, const CoattArgs& a, const int32_t* live);
bool try_coatt_bwd_lean(cudaStream_t st, const Dims& dm, const CoattBwdArgs& a, const int32_t* live);
int coatt_bwd_num_ctas();
void launch_coatt_bwd(cudaStream_t st, const Dims& dm, const CoattBwdArgs& a);

struct TargetBwdArgs {
    const int32_t* length;
    const float* w_item; const float* w_user;
    const float* q0; const float* dq0;         // [B, Ds]
    const float* dfc_in; int fc_off; int ldfc; // d fc_in[b, fc_off:]
    const float* sdz;
    float* grad_rows;
    float* partials; int n_partials;           // [n_partials, Di + 1 + Du + 1]: dWt_item | db_item | dWt_user | db_user
};
int target_bwd_num_ctas();
void launch_target_bwd(cudaStream_t st, const Dims& dm, const TargetBwdArgs& a);

// reduce the co-attention partials of both kernels into the flat dense gradient buffer (split 0)
void launch_coatt_grad_reduce(cudaStream_t st, const Dims& dm, const float* coatt_partials, int n_coatt,
                              const float* target_partials, int n_target,
                              float* g_w_item, float* g_b_item, float* g_w_user, float* g_b_user);

// ---------------------------------------------------------------- sequence encoder + head (seq.cu)
struct GruArgs {
    const int32_t* length;
    const float* px[2];            // [M, 3H] input projections (no bias)
    const float* wg[2]; const float* bg[2];   // gates kernel [ldx, 2H] (rows Ds.. are the state part), bias [2H]
    const float* wc[2]; const float* bc[2];   // candidate kernel [ldx, H], bias [H]
    float* xhg[2]; float* xhc[2];  // [M, ldx]: state part (cols Ds..) written: h_prev and r*h_prev
    float* r[2]; float* u[2]; float* c[2];    // [M, H] saved gate values
    float* out; int ldout;         // key: side s writes out[m, s*H : (s+1)*H]  (0 for t >= length)
    float* last;  int ldlast;      // optional final state -> last[b, s*H:...] (RIA), may be null
};
void launch_gru_fwd(cudaStream_t st, const Dims& dm, const GruArgs& a);

struct GruBwdArgs {
    const int32_t* length;
    const float* wg[2]; const float* wc[2];
    const float* xhg[2];
    const float* r[2]; const float* u[2]; const float* c[2];
    const float* dout; int lddout;   // d key[:, s*H:(s+1)*H]
    const float* dlast; int lddlast; // optional gradient of the final state (RIA), may be null
    float* dpx[2];                   // [M, 3H]
};
void launch_gru_bwd(cudaStream_t st, const Dims& dm, const GruBwdArgs& a);

// batch-norm in inference mode (score.py:69): z = x * gamma/sqrt(var+eps) + (beta - mean*inv)
void launch_bn_fwd(cudaStream_t st, int B, int F, const float* x, const float* gamma, const float* beta,
                   const float* mean, const float* var, float* z);
void launch_bn_bwd(cudaStream_t st, int B, int F, const float* x, const float* dz, const float* gamma,
                   const float* mean, const float* var, float* dx, float* dgamma, float* dbeta);

// fused prediction head (chain.cu): bn1 -> fc1 -> fc2 -> fc3 -> sigmoid -> per-sample loss and d loss / d logit
struct FcArgs {
    int B, F;
    const float* fc_in; const float* gamma; const float* beta; const float* mean; const float* var;
    const float* w1; const float* b1; const float* w2; const float* b2; const float* w3; const float* b3;
    const int32_t* label; const Hyper* hp;
    float* z0; float* g1; float* g2; float* y; float* loss_b; float* dlogit;
};
void launch_fc_fwd(cudaStream_t st, const FcArgs& a);
struct FcBwdArgs {
    int B, F;
    const float* dlogit; const float* g2; const float* g1;
    const float* w3; const float* w2; const float* w1; const float* gamma; const float* var;
    const float* w2t; const float* w1t;   // derived transposes: fc2^T [80,200], fc1^T [200,F] (prep_weights)
    const Hyper* hp;
    float* dg2; float* dg1; float* dz0; float* dfc_in;
};
void launch_fc_bwd(cudaStream_t st, const FcBwdArgs& a);
// dgamma / dbeta of the inference-mode batch norm (the dx part is produced by fc_bwd)
void launch_bn_param_grads(cudaStream_t st, int B, int F, const float* x, const float* dz, const float* mean,
                           const float* var, float* dgamma, float* dbeta, int splits, int64_t split_stride);

// logit = g2 . w3 + b3; y = sigmoid; per-sample log-loss (eps 1e-7) and d loss / d logit
void launch_head(cudaStream_t st, int B, int F, const float* g2, const float* w3, const float* b3,
                 const int32_t* label, const Hyper* hp, float* y, float* loss_b, float* dlogit);
// loss = sum_b loss_b * inv_batch + reg_lambda * l2sum    (fixed-order reduction)
// early (may be null; device-visible pinned HOST memory): 4 floats {loss, L2 part, *err_flag as bits, hp->seq as bits},
// the sequence number last and behind a system fence - the result packet the host polls for, available as soon as the
// forward pass is done
void launch_loss_final(cudaStream_t st, int B, const float* loss_b, const float* l2sum, const Hyper* hp, float* loss,
                       const int32_t* err_flag = nullptr, float* early = nullptr);


// ---------------------------------------------------------------- fused attention block (attn.cu)
// Derived weights rebuilt once per step from the flat parameter buffer P into D:
//   D[dst_off + r*ld_dst + c] = P[a_off + r*a_rs + c*a_cs] (+/-) P[b_off + r*a_rs + c*a_cs]     (b_off < 0: no second term)
struct PrepOp { int64_t dst_off, a_off, b_off; int ld_dst, rows, cols, a_rs, a_cs, sign; };
struct PrepOps { PrepOp op[24]; int n; };
void launch_prep_weights(cudaStream_t st, const PrepOps& ops, const float* P, float* D);

struct AttQArgs {
    int B, Ds, Dk;
    const float* q0;                       // [B, Ds] = [target_user || target_item]
    const float* wq; const float* bq;      // query projection [Ds, Dk] (score.py:172)
    const float* Wac; const float* b1;     // (Wa + Wc) [Dk, 80], bias of the first attention layer
    float* q; float* U;                    // [B, Dk], [B, 80]
};
void launch_att_q(cudaStream_t st, const AttQArgs& a);

struct AttFwd2Args {
    int B, T, Dk, H, G;                    // G (samples per CTA) is set by the launcher
    const int32_t* length;
    const float* q; const float* U; const float* key;
    const float* W1e;                      // [2*Dk, 80]: rows (Wb - Wc) | Wd
    const float* w2; const float* b2; const float* w3; const float* b3;
    float* qk;                             // [M, Dk] q*key (operand of the Wd weight gradient)
    float* f1; float* f2;                  // [M, 80], [M, 40]
    float* score;                          // [B, T]
    float* fc_in; int ldfc; int model_type;
};
void launch_att_fwd2(cudaStream_t st, AttFwd2Args a);

struct AttBwd2Args {
    int B, T, Dk, H, G;
    const int32_t* length;
    const float* q; const float* key; const float* f1; const float* f2; const float* score;
    const float* dfc_in; int ldfc; int model_type;
    const float* w3;
    const float* W2T;                      // [40, 80]
    const float* W1eT;                     // [80, 2*Dk]: columns (Wb - Wc)^T | Wd^T
    float* ds; float* df2; float* df1;     // [M], [M, 40], [M, 80]
    float* dkey;                           // [M, Dk]
    float* sdf1;                           // [B, 80]  sum_t d f1
    float* dqD;                            // [B, Dk]  sum_t dD * key
};
void launch_att_bwd2(cudaStream_t st, AttBwd2Args a);

struct AttQbArgs {
    int B, Ds, Dk;
    const float* sdf1; const float* dqD;
    const float* WacT;                     // [80, Dk]
    const float* WqT;                      // [Dk, Ds]
    float* dq; float* dq0;                 // [B, Dk], [B, Ds]
};
void launch_att_qb(cudaStream_t st, const AttQbArgs& a);

// out[m, :N] = X[m, :K] W[:K, :N] (+ bias); K, N, ldx, ldw, ldo multiples of 4, 16-byte aligned bases; up to 4 problems
struct RowGemmArgs { const float* X; int ldx; const float* W; int ldw; const float* bias; float* out; int ldo; int M, K, N; };
struct RowGemmBatch { RowGemmArgs p[4]; };
void launch_rowgemm(cudaStream_t st, const RowGemmArgs* list, int n);

// ---------------------------------------------------------------- optimizer + scatter (scatter.cu)
// partial sums of v*v/2 over L2-regularised dense parameters (flags bit0): out[L2_PARTS]
constexpr int L2_PARTS = 32;
void launch_l2_sum(cudaStream_t st, const float* params, const uint8_t* flags, int n, float* out);
// G[i] = sum_s partial[s][i]
// optional derived range afterwards: g[dst + i] = g[a + i] - g[b + i], i < count  (dWc = dWa - dWb of the attention's first layer)
// optional fused dense Adam step on element i (same arithmetic as launch_dense_adam)
struct DenseAdamArgs { float* p; float* m; float* v; const uint8_t* flags; const Hyper* hp; float* alpha_hist; };
void launch_reduce_partials(cudaStream_t st, const float* partials, int splits, int n, float* g,
                            int64_t derive_dst = -1, int64_t derive_a = 0, int64_t derive_b = 0, int derive_count = 0,
                            const DenseAdamArgs* adam = nullptr);
void launch_add_l2(cudaStream_t st, float* g, const float* p, const uint8_t* flags, int n, const Hyper* hp);
// dense Adam on the flat parameter buffer: g = G + reg*p (flags bit0), skip non-trainables (flags bit1 clear)
void launch_dense_adam(cudaStream_t st, float* p, float* m, float* v, const float* g, const uint8_t* flags,
                       int n, const Hyper* hp, float* alpha_hist /* LAZY: alpha_hist[step] = alpha, may be null */);

struct SortBufs {
    int32_t* keys[2]; int32_t* vals[2];   // ping-pong
    uint32_t* hist;                        // [256 * nblocks]
    int32_t* runs;                         // [8 * n_cap] short-run descriptors of the sorted list (launch_emb_runs)
    int32_t* runs_long;                    // [4 * emb_runs_long_cap(n_cap)] medium / long run descriptors
    // runs longer than RUN_CHUNK entries are reduced by several CTAs (one chunk each) and combined in chunk order:
    float* part;                           // [emb_runs_part_cap(n_cap) * d] partial sums, one row per chunk
    int32_t* slotinfo;                     // [4 * emb_runs_part_cap(n_cap)] per chunk {first chunk slot of its run, chunks, run start, 0}
    int32_t* done;                         // [emb_runs_part_cap(n_cap)] chunks finished per run (zero between launches)
    int n_cap;
};
size_t sort_hist_elems(int64_t n);
// stable LSD radix sort of (key, position) pairs; keys_in is left untouched, values start as 0..n-1;
// returns the index (0/1) of the ping-pong buffer that holds the result
int launch_sort_pairs(cudaStream_t st, SortBufs& sb, const int32_t* keys_in, int64_t n, int key_bits);
// the same sort in two instalments (the step runs the first pass before the gather kernel and the rest after it)
int sort_num_passes(int key_bits);
int launch_sort_passes(cudaStream_t st, SortBufs& sb, const int32_t* keys_in, int64_t n, int p_begin, int p_end);

// Run descriptors of the sorted (key, position) list, three tiers by run length (scatter.cu):
//   runs       8 int32 per run of <= 4 entries: {key, start, n, pos0, pos1, pos2, pos3, 0}           (capacity n runs)
//   runs_long  4 int32 per longer run {key, start, count, 0}: runs of up to RUN_M (32, scatter.cu) entries from the front, longer ones from
//              the back of a buffer of emb_runs_long_cap(n) descriptors
//   counters   4 int32: number of short / medium / long runs, and of all runs (= unique rows)
//   counters   8 int32: [0..2] number of short / medium / long descriptors, [3] all runs (= unique rows), [4] chunk slots used
int64_t emb_runs_long_cap(int64_t n);
int64_t emb_runs_part_cap(int64_t n);
void launch_emb_runs(cudaStream_t st, const int32_t* skeys, const int32_t* spos, int64_t n, int32_t* runs,
                     int32_t* runs_long, int32_t* counters, int32_t* slotinfo);
struct EmbUpdateArgs {
    const int32_t* skeys; const int32_t* spos; int64_t n;
    const int32_t* runs; const int32_t* runs_long; int64_t long_cap; const int32_t* counters;   // from launch_emb_runs
    float* part; const int32_t* slotinfo; int32_t* done;                                       // chunked long runs (SortBufs)
    const float* grad_rows; int d;
    float* emb; float* m; float* v; int es; int32_t* last_step;   // es: floats between consecutive rows of emb / m / v
    const float* alpha_hist;       // LAZY: rows that are not current through step-1 are replayed first (may be null)
    const Hyper* hp;
    int mode;                      // 0: apply Adam; 1 / 2: export the run sums to out_rows / out_heads at the run's first sorted
                                   // index / at head_slot[first sorted index] (compact, ascending by key), no update
    float* out_rows; int32_t* out_heads;
    const int32_t* head_slot; int64_t out_cap;   // mode 2 (launch_head_slots); out_cap = capacity of the compact list
};
void launch_emb_update(cudaStream_t st, const EmbUpdateArgs& a);

// ---- data-parallel exchange (replicated table): every rank contributes ONE packed block of int32 / float words
//   [0,128)                      header: unique-row count, step sequence number, loss (total, L2 part) as float bits
//   [dense_off, +n_dense)        this rank's dense gradient (flat buffer, 1/global_batch scaling applied)
//   [keys_off, +cap)             unique row ids, ascending, zeros past the count
//   [keys_off+cap, +cap*d)       one gradient row per id
// blocks of all ranks lie `stride` words apart in the gathered buffer `base`; every offset is a multiple of 128 words
// (rows are 16-byte aligned for any d).
struct DpLayout { const int32_t* base; int world; int d; int64_t stride, dense_off, keys_off, cap; };
// destinations of a peer-memory push: where this rank's block goes in every replica's gathered buffer
constexpr int DP_MAX_WORLD = 16;
struct DpPeers { int32_t* dst[DP_MAX_WORLD]; int world; };
void launch_dp_push(cudaStream_t st, const int32_t* block, int64_t words, const DpPeers& peers);
int64_t head_slot_tiles(int64_t n);
// head_slot[i] = number of run heads before sorted index i, written at run heads only; tile_counts: head_slot_tiles(n) ints
void launch_head_slots(cudaStream_t st, const int32_t* skeys, int64_t n, int32_t* tile_counts, int32_t* head_slot);
void launch_dp_count(cudaStream_t st, const int32_t* counters, const Hyper* hp, int32_t* slot);
// header + dense gradient + zero padding of the id list of this rank's block (the rows come from launch_emb_update, mode 2)
void launch_dp_pack_misc(cudaStream_t st, int32_t* block, const DpLayout& L, const int32_t* counters, const Hyper* hp,
                         const float* loss_dev, const float* g, int n_dense);
// embedding half of the data-parallel finish over the gathered blocks, one kernel: per id, the ranks' rows added in
// rank order + the row's Adam step (a: emb / m / v / last_step / alpha_hist / hp / d as for launch_emb_update)
// sample: dp_sample_count(world, cap) ints of scratch (sampled copies of the id lists, built by the same call)
int64_t dp_sample_count(int world, int64_t cap);
void launch_dp_apply(cudaStream_t st, const DpLayout& L, const EmbUpdateArgs& a, int32_t* sample, int32_t* err_flag);
// rank-ordered sum of the gathered dense gradients + dense Adam; loss_out[0] = global loss (double)
void launch_dp_dense_adam(cudaStream_t st, const DpLayout& L, float* p, float* m, float* v, float* g_out,
                          const uint8_t* flags, int n, const Hyper* hp, float* alpha_hist, double* loss_out);

// DENSE mode: zero-gradient Adam step for every row whose last_step != hp->step
void launch_emb_dense_sweep(cudaStream_t st, float* emb, float* m, float* v, int32_t* last_step, int64_t V, int d, int es,
                            const Hyper* hp);
// LAZY mode: replay skipped zero-gradient steps of the rows about to be gathered (keys in position order;
// duplicates are resolved by an atomic claim on last_step, the replay result does not depend on the winner).
// claim_list: 2*n int32 of scratch, claim_counter: one int32.
void launch_emb_catchup_rows(cudaStream_t st, const int32_t* keys, int64_t n, int64_t V, float* emb, float* m, float* v,
                             int32_t* last_step, int d, int es, const float* alpha_hist, const Hyper* hp, int32_t* claim_list,
                             int32_t* claim_counter);
// LAZY mode: bring the whole table up to `upto_step` (before read-back / save / eval of everything)
void launch_emb_catchup_all(cudaStream_t st, float* emb, float* m, float* v, int32_t* last_step, int64_t V, int d, int es,
                            const float* alpha_hist, int upto_step);

// ---- row-sharded table (shard.cu): positions grouped by owner = id % world (stable), see score_shard_plan
//   owner [n] scratch; counts [world + 1] (last: dummy positions); send_rows / sel [n]; mini_keys [n]
void launch_shard_plan(cudaStream_t st, SortBufs& sb, const int32_t* keys, int64_t n, int world, int32_t* owner,
                       int32_t* counts, int32_t* send_rows, int32_t* sel, int32_t* mini_keys);
void launch_shard_pack_grads(cudaStream_t st, const float* grad_rows, const int32_t* sel, const int32_t* counts, int world,
                             int64_t n, int d, float* out);

// peer-memory variant of the row-sharded exchange (shard.cu): rows / gradient rows are stored straight into the peer's
// staged table / gradient buffer; p[r] = that buffer of rank r as mapped into this process
constexpr int SHARD_MAX_PEERS = 64;
struct ShardPeers { float* p[SHARD_MAX_PEERS]; };
void launch_shard_serve_push(cudaStream_t st, const float* table, int es, int d, int64_t V, const int32_t* want, int64_t n_recv,
                             const int32_t* cm, int world, int me, const ShardPeers& peers, int32_t* err_flag);
void launch_shard_grad_push(cudaStream_t st, const float* grad_rows, const int32_t* sel, int64_t n, int d, const int32_t* cm,
                            int world, int me, const ShardPeers& peers);

// out[i] = table[idx[i]] (idx 0 -> zeros; out-of-range -> zeros + error flag): owner side of a sharded gather
void launch_gather_rows(cudaStream_t st, const float* table, int es, const int32_t* idx, int64_t n, int d, int64_t V, float* out,
                        int32_t* err_flag);

// device-side TF-default initialisers
// row_len / row_stride: element i goes to p[(i / row_len) * row_stride + i % row_len] (0: contiguous)
void launch_init_trunc_normal(cudaStream_t st, float* p, int64_t n, uint64_t seed, uint32_t stream_id, int row_len = 0,
                              int row_stride = 0);
void launch_init_uniform(cudaStream_t st, float* p, int64_t n, float limit, uint64_t seed, uint32_t stream_id);
void launch_fill(cudaStream_t st, float* p, int64_t n, float v);
void launch_fill_i32(cudaStream_t st, int32_t* p, int64_t n, int32_t v);

// ---------------------------------------------------------------- eval metrics (metrics.cu)
// sums[0] = sum of log-loss terms; sums[2..7] = sums of ndcg5 ndcg10 hr1 hr5 hr10 mrr over groups;
// sums[8] = sum over positives of mid-ranks, sums[9] = number of positives.  terms: >= max(2n, 6n/group) doubles.
cudaError_t compute_eval_metrics(cudaStream_t st, const float* preds, const int32_t* iids, const int32_t* labels,
                                 int64_t n, int group, SortBufs& sb, int32_t* keys, int32_t* pos_rank, double* terms,
                                 double* sums);

}  // namespace score
