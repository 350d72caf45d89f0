// Model handle, workspace and step orchestration behind the C ABI (include/score_b200.h).
//
// One handle = one model on one device: the single shared embedding table with its Adam slots,
// one flat buffer of every dense variable (TF names and creation order of score.py), the
// per-batch workspace, two streams (main + scatter/sort branch) and, optionally, one captured
// CUDA graph per batch size.  The step is the data-parallel hot path of
// code/score/score.py:101-133 (train / eval) with the graph of score.py:188-224 underneath.
#include "../../include/score_b200.h"
#include "kernels.h"

#include <math.h>
#include <stdio.h>
#include <string.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

using namespace score;

namespace {

thread_local std::string g_create_error;

constexpr int kSplits = 32;   // fixed number of batch-row chunks for weight gradients (deterministic two-stage sum)

struct Tensor {
    std::string name;
    int64_t rows, cols;
    int64_t off;       // float offset in the flat dense buffer (unused for emb_mtx)
    bool is_emb;
    uint8_t flags;     // bit0: L2-regularised (score.py:92-94), bit1: trainable
};

struct Buf {
    void* ptr; size_t count; int dtype;   // 0 float32, 1 int32
};

enum Probe { PR_COATT_FWD = 0, PR_COATT_BWD, PR_EMB_UPDATE, PR_SORT, PR_STEP, PR_FWD_DENSE, PR_BWD_DENSE, PR_CATCHUP, PR_COUNT };

}  // namespace

struct ScoreModel {
    ScoreConfig cfg;
    int device = 0;
    std::string err;
    Dims dm;                        // dm.B / offsets are per call
    cudaStream_t st = nullptr, st2 = nullptr, st_w = nullptr;   // main, sort branch, weight-gradient branch
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_l2 = nullptr, ev_w = nullptr, ev_tgt = nullptr, ev_q = nullptr, ev_prep = nullptr;
    cudaEvent_t ev_pool[16] = {nullptr}; int ev_next = 0;
    bool side_w = false;   // a step is being enqueued with the weight-gradient branch forked

    // parameters
    std::vector<Tensor> tensors;
    std::map<std::string, int> tindex;
    int64_t n_dense = 0;
    // Embedding state: ONE buffer of V records [var | m | v] (3*d floats each), so the optimizer kernels touch one
    // contiguous 12*d-byte run per row instead of three rows in three arrays (a d=16 row is half a 128-byte DRAM
    // access: profiles/README.md).  emb / emb_m / emb_v point at the three fields of record 0; `es` = 3*d is the row stride.
    float *emb_tab = nullptr, *emb = nullptr, *emb_m = nullptr, *emb_v = nullptr; int es = 0;
    int es_fwd = 0;   // row stride of emb_fwd (es, or d for a staged mini-table)
    int32_t* last_step = nullptr;   // per-row step at which the row is current (DENSE / LAZY)
    float *P = nullptr, *G = nullptr, *M1 = nullptr, *V1 = nullptr, *PG = nullptr;
    uint8_t* flags = nullptr;
    // derived weights of the fused dense chains (attn.cu, chain.cu), rebuilt from P at the start of every step
    float* Dv = nullptr; int64_t n_derived = 0; PrepOps prep{};
    int64_t dv_W1e = 0, dv_Wac = 0, dv_W1eT = 0, dv_WacT = 0, dv_W2T = 0, dv_WqT = 0, dv_fc1T = 0, dv_fc2T = 0;
    int64_t dv_Wx[2] = {0, 0}, dv_WxT[2] = {0, 0};   // GRU input-side kernels [Wg_x | Wc_x] ([Ds,3H]) and their transpose
    // LAZY Adam: alpha of every step since the last re-base, indexed by the ABSOLUTE step through a shifted pointer
    // (alpha_hist = alpha_buf - hist_base; valid for steps in (hist_base, hist_base + alpha_cap)).  When the window is
    // nearly full every row is brought up to date and the window restarts at the current step (lazy_rebase).
    float* alpha_hist = nullptr; float* alpha_buf = nullptr; int64_t alpha_cap = 0; int64_t hist_base = 0;

    // optimizer scalars (fp32 like the TF slot variables)
    int32_t step = 0;
    int32_t sample_base = 0;   // score_set_sample_offset: global index of the local batch's first sample
    float beta1_power = 0.9f, beta2_power = 0.999f;

    // workspace
    int cap_B = 0;
    std::vector<void*> allocs;       // per-capacity allocations (freed on regrow)
    std::map<std::string, Buf> bufs;
    int32_t *ids = nullptr, *label = nullptr, *length = nullptr, *keys = nullptr;
    int32_t* live = nullptr;   // live-slice work list of the lean co-attention kernels (launch_build_keys)
    float *q0, *c_item, *c_user, *xhg[2], *xhc[2], *key, *save_r, *save_w, *px[2], *gr[2], *gu[2], *gc[2];
    float *q, *attU, *qk, *f1, *f2, *score, *fc_in, *z0, *g1, *g2, *y, *loss_b, *dlogit;
    float *dg2, *dg1, *dz0, *dfc_in, *ds, *df2, *df1, *sdf1, *dqD, *dkey, *dq, *dq0, *dpx[2], *dx[2], *sdz, *grad_rows;
    float *coatt_part, *target_part;
    int n_coatt_part = 0, n_target_part = 0;
    SortBufs sb{};
    int sort_out = 0;
    float *seg_rows = nullptr; int32_t* seg_heads = nullptr; int64_t seg_cap = 0;
    int64_t last_N = 0;
    int32_t* claim_list = nullptr; int32_t* claim_ext = nullptr; int64_t claim_ext_cap = 0;   // LAZY catch-up scratch
    int32_t* claim_counter = nullptr;
    int32_t* n_heads_dev = nullptr;   // number of run heads in the sorted key list (emb_heads_kernel)
    float *l2sum = nullptr, *loss_dev = nullptr;   // loss_dev[0] = total loss, loss_dev[1] = reg_lambda * l2 part
    int32_t* err_flag = nullptr;
    // Per-step parameters the kernels read from device memory (one H2D copy per step; captured graphs stay valid):
    // the hyper-parameters and the pointer table of the current batch.
    struct StepParams { Hyper hp; BatchPtrs bp; };
    StepParams* step_dev = nullptr;
    Hyper* hyper_dev = nullptr;        // &step_dev->hp
    BatchPtrs* bp_dev = nullptr;       // &step_dev->bp
    BatchPtrs bp_cur{};                // pointer table of the batch uploaded last (host copy)
    // pinned host staging.  hyper_ring: one slot per in-flight step so an asynchronous caller never overwrites a
    // struct whose H2D copy has not executed yet; hyper_host points at the slot of the current call.
    static constexpr int kHyperSlots = 32;
    StepParams* hyper_ring = nullptr; cudaEvent_t hyper_ev[kHyperSlots] = {nullptr}; bool hyper_used[kHyperSlots] = {false};
    int hyper_next = 0;
    int32_t hyper_seq = 0;
    Hyper* hyper_host = nullptr;
    float* loss_host = nullptr;
    int32_t* err_host = nullptr;

    // table / key list the forward and backward kernels read: the handle's own, or a staged per-batch
    // mini-table when the embedding rows live on other ranks (row-sharded tables)
    const float* emb_fwd = nullptr; const int32_t* keys_fwd = nullptr;
    // sort buffers for externally supplied (key, gradient row) lists (multi-GPU finish)
    SortBufs sb_ext{}; int64_t sb_ext_cap = 0;
    bool local_sorted = false;   // score_step_begin sorted this rank's own keys (side stream)
    std::map<int, cudaGraphExec_t> graphs_begin; std::map<int, int64_t> graph_kernels_begin; std::map<int, int> warm_begin;
    cudaEvent_t ev_keys = nullptr, ev_fc = nullptr, ev_att = nullptr, ev_qb = nullptr;
    bool sort_deferred = false;   // enqueue_forward forks the sort branch right after the gather kernel
    bool sort_dp = false;         // the sort branch also publishes the unique-row count and the run ranks (data-parallel)
    bool begun = false; float begun_lr = 0.f;
    const int32_t* last_sorted = nullptr; int64_t last_sorted_n = 0;   // sorted key list of the last optimizer step
    // data-parallel packed exchange (score_dp_*): rank of every run head (compact sorted export), the early unique-row
    // count for the host (its own stream + event, so reading it never drains the main stream), this rank's block
    int32_t *head_slot = nullptr, *hs_tiles = nullptr;          // workspace (per capacity)
    int32_t *cnt_slot = nullptr, *cnt_host = nullptr;           // {count, step sequence number}: device / pinned host
    cudaStream_t st_cnt = nullptr; cudaEvent_t ev_counts = nullptr;
    int32_t begin_seq = -1; bool warned_stale = false;
    int32_t* dp_block = nullptr; int64_t dp_block_words = 0;
    int32_t* dp_sample = nullptr; int64_t dp_sample_cap = 0;   // sampled id lists of the gathered blocks (launch_dp_apply)
    double *loss_glob = nullptr, *loss_glob_host = nullptr;
    // Synchronous train step without draining the device: the loss (+ error flag) leaves the device right after the
    // forward pass (result packet written by loss_final into pinned host memory, which the host polls); backward and
    // update keep running while the caller prepares the next batch.  Host batches are staged through two buffers on a copy stream, so the next step's H2D overlaps them.
    float* early_host = nullptr; bool early_on = false;
    int32_t *ids_stage[2] = {nullptr, nullptr}, *lab_stage[2] = {nullptr, nullptr}, *len_stage[2] = {nullptr, nullptr};
    cudaStream_t st_h2d = nullptr; cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_stage_free[2] = {nullptr, nullptr};
    bool stage_busy[2] = {false, false}; int stage_idx = 0; int stage_last = -1;

    // row-sharded exchange plan (score_shard_plan): buffers sized for sh_cap positions
    int64_t sh_cap = 0; int sh_world = 0;
    int32_t *sh_owner = nullptr, *sh_counts = nullptr, *sh_send_rows = nullptr, *sh_sel = nullptr, *sh_mini = nullptr;
    float *sh_staged = nullptr, *sh_grad_send = nullptr;
    const int32_t* presorted_keys = nullptr; int64_t presorted_n = 0; int presorted_out = 0; cudaEvent_t ev_presort = nullptr;
    int32_t* sh_counts_host = nullptr; cudaEvent_t ev_sh_counts = nullptr;   // count matrix read-back (pinned, own event)
    const float* sh_ext_staged[2] = {nullptr, nullptr};   // caller-owned staged tables with stable addresses (graph capture)

    // graphs
    std::map<int, cudaGraphExec_t> graphs_train;
    std::map<int, int> warm_train;
    std::map<int, int64_t> graph_kernels;   // kernels inside the captured graph of each batch size
    bool capturing = false;

    // timing probes
    bool probes_on = false;
    cudaEvent_t pr_beg[PR_COUNT], pr_end[PR_COUNT];
    double pr_ms[PR_COUNT]; int64_t pr_n[PR_COUNT];
    bool pending_step = false;
    int last_mode = -1;

    int64_t launches0 = 0;
};

namespace {

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char b_[512];                                                                          \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            h->err = b_;                                                                           \
            return SCORE_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

int fail(ScoreModel* h, int code, const std::string& msg) {
    h->err = msg;
    return code;
}

const Tensor* find_tensor(ScoreModel* h, const std::string& name) {
    auto it = h->tindex.find(name);
    return it == h->tindex.end() ? nullptr : &h->tensors[it->second];
}
float* pp(ScoreModel* h, const char* name) { return h->P + find_tensor(h, name)->off; }
int64_t po(ScoreModel* h, const char* name) { return find_tensor(h, name)->off; }

// TF variable names by role (oracle/score_ref.py:role_names states the numbering rule)
struct Names {
    std::string co_item, co_user, att_q, att1, att2, att3;
};
Names role_names(int model_type) {
    Names n;
    int k = 0;
    auto next = [&]() { std::string s = k == 0 ? "dense" : "dense_" + std::to_string(k); ++k; return s; };
    if (model_type != SCORE_MODEL_RCA && model_type != SCORE_MODEL_RRN) { n.co_item = next(); n.co_user = next(); }
    if (model_type != SCORE_MODEL_RIA && model_type != SCORE_MODEL_RRN) { n.att_q = next(); n.att1 = next(); n.att2 = next(); n.att3 = next(); }
    return n;
}

void build_registry(ScoreModel* h) {
    const ScoreConfig& c = h->cfg;
    const int d = c.eb_dim, H = c.hidden_size, K = c.obj_per_time_slice;
    const int Du = c.user_fnum * d, Di = c.item_fnum * d, Ds = Du + Di;
    const int mt = c.model_type;
    const bool rrn = mt == SCORE_MODEL_RRN;
    const bool has_coatt = mt != SCORE_MODEL_RCA && !rrn, has_att = mt != SCORE_MODEL_RIA && !rrn;
    const int Dk = has_coatt ? 2 * H + 4 * K : 2 * H;
    const int nstate = (mt == SCORE_MODEL_SCORE_USER || mt == SCORE_MODEL_SCORE_ITEM) ? 1 : 2;
    const int Dfc = nstate * H + Du + Di;
    Dims& dm = h->dm;
    dm.V = c.feature_size; dm.T = c.max_time_len; dm.K = K; dm.d = d; dm.H = H;
    dm.fu = c.user_fnum; dm.fi = c.item_fnum; dm.Du = Du; dm.Di = Di; dm.Ds = Ds; dm.Dk = Dk; dm.Dfc = Dfc;
    dm.ldx = Ds + H; dm.model_type = mt; dm.B = 0;
    // RRN (slice_model.py:158-159): user side = reduce_sum(user_1hop) [Di], item side = reduce_sum(item_1hop) [Du]
    dm.Dx[0] = rrn ? Di : Ds; dm.Dx[1] = rrn ? Du : Ds;
    dm.ldxs[0] = dm.Dx[0] + H; dm.ldxs[1] = dm.Dx[1] + H; dm.hop1_only = rrn ? 1 : 0;

    int64_t off = 0;
    auto add = [&](const std::string& name, int64_t r, int64_t cdim, uint8_t flags, bool emb = false) {
        Tensor t{name, r, cdim, emb ? -1 : off, emb, flags};
        if (!emb) off += (r * cdim + 3) / 4 * 4;
        h->tindex[name] = (int)h->tensors.size();
        h->tensors.push_back(t);
    };
    auto kb = [&](const std::string& prefix, int64_t r, int64_t cdim) {
        add(prefix + "/kernel", r, cdim, 3);
        add(prefix + "/bias", 1, cdim, 2);
    };
    Names nm = role_names(mt);
    add("emb_mtx", c.feature_size, d, 2, true);
    if (has_coatt) { kb(nm.co_item, 3 * Di, 1); kb(nm.co_user, 3 * Du, 1); }
    {
        const char* gsides[2] = {"gru_user_side", "gru_item_side"};
        for (int sd = 0; sd < 2; ++sd) {
            kb(std::string(gsides[sd]) + "/gru_cell/gates", dm.Dx[sd] + H, 2 * H);
            kb(std::string(gsides[sd]) + "/gru_cell/candidate", dm.Dx[sd] + H, H);
        }
    }
    if (has_att) { kb(nm.att_q, Ds, Dk); kb(nm.att1, 4 * Dk, 80); kb(nm.att2, 80, 40); kb(nm.att3, 40, 1); }
    add("bn1/gamma", 1, Dfc, 3);
    add("bn1/beta", 1, Dfc, 3);
    add("bn1/moving_mean", 1, Dfc, 0);
    add("bn1/moving_variance", 1, Dfc, 0);
    kb("fc1", Dfc, 200); kb("fc2", 200, 80); kb("fc3", 80, 1);
    h->n_dense = off;

    // derived weights: combinations / transposes the fused chains stream (see attn.cu header)
    int64_t doff = 0;
    PrepOps& po_ = h->prep;
    po_.n = 0;
    auto reserve = [&](int64_t n) { int64_t o = doff; doff += (n + 3) / 4 * 4; return o; };
    auto op = [&](int64_t dst, int ld, int rows, int cols, int64_t a, int64_t b, int a_rs, int a_cs, int sign) {
        PrepOp& o = po_.op[po_.n++];
        o.dst_off = dst; o.ld_dst = ld; o.rows = rows; o.cols = cols; o.a_off = a; o.b_off = b; o.a_rs = a_rs; o.a_cs = a_cs; o.sign = sign;
    };
    auto toff = [&](const std::string& name) { return h->tensors[h->tindex[name]].off; };
    if (has_att) {
        const int64_t w1 = toff(nm.att1 + "/kernel"), blk = (int64_t)Dk * 80;   // row blocks Wa | Wb | Wc | Wd
        h->dv_W1e = reserve(2 * blk); h->dv_Wac = reserve(blk); h->dv_W1eT = reserve(2 * blk); h->dv_WacT = reserve(blk);
        h->dv_W2T = reserve(40 * 80); h->dv_WqT = reserve((int64_t)Dk * Ds);
        op(h->dv_W1e, 80, Dk, 80, w1 + blk, w1 + 2 * blk, 80, 1, -1);              // Wb - Wc
        op(h->dv_W1e + blk, 80, Dk, 80, w1 + 3 * blk, -1, 80, 1, 0);               // Wd
        op(h->dv_Wac, 80, Dk, 80, w1, w1 + 2 * blk, 80, 1, +1);                    // Wa + Wc
        op(h->dv_W1eT, 2 * Dk, 80, Dk, w1 + blk, w1 + 2 * blk, 1, 80, -1);         // (Wb - Wc)^T
        op(h->dv_W1eT + Dk, 2 * Dk, 80, Dk, w1 + 3 * blk, -1, 1, 80, 0);           // Wd^T
        op(h->dv_WacT, Dk, 80, Dk, w1, w1 + 2 * blk, 1, 80, +1);                   // (Wa + Wc)^T
        op(h->dv_W2T, 80, 40, 80, toff(nm.att2 + "/kernel"), -1, 1, 40, 0);        // W2^T
        op(h->dv_WqT, Ds, Dk, Ds, toff(nm.att_q + "/kernel"), -1, 1, Dk, 0);       // Wq^T
    }
    {
        const char* sides[2] = {"gru_user_side", "gru_item_side"};
        for (int sd = 0; sd < 2; ++sd) {
            const int64_t wg = toff(std::string(sides[sd]) + "/gru_cell/gates/kernel");        // [Ds+H, 2H]
            const int64_t wc = toff(std::string(sides[sd]) + "/gru_cell/candidate/kernel");    // [Ds+H, H]
            const int Dx = dm.Dx[sd];
            h->dv_Wx[sd] = reserve((int64_t)Dx * 3 * H); h->dv_WxT[sd] = reserve((int64_t)3 * H * Dx);
            op(h->dv_Wx[sd], 3 * H, Dx, 2 * H, wg, -1, 2 * H, 1, 0);
            op(h->dv_Wx[sd] + 2 * H, 3 * H, Dx, H, wc, -1, H, 1, 0);
            op(h->dv_WxT[sd], Dx, 2 * H, Dx, wg, -1, 1, 2 * H, 0);
            op(h->dv_WxT[sd] + (int64_t)2 * H * Dx, Dx, H, Dx, wc, -1, 1, H, 0);
        }
    }
    h->dv_fc1T = reserve((int64_t)200 * Dfc); h->dv_fc2T = reserve(80 * 200);
    op(h->dv_fc1T, Dfc, 200, Dfc, toff("fc1/kernel"), -1, 1, 200, 0);              // fc1^T
    op(h->dv_fc2T, 200, 80, 200, toff("fc2/kernel"), -1, 1, 80, 0);                // fc2^T
    h->n_derived = doff;
}

int alloc_params(ScoreModel* h) {
    const int64_t V = h->cfg.feature_size, d = h->cfg.eb_dim;
    h->es = 3 * d;
    CK(cudaMalloc(&h->emb_tab, sizeof(float) * V * h->es));
    CK(cudaMemsetAsync(h->emb_tab, 0, sizeof(float) * V * h->es, h->st));
    h->emb = h->emb_tab; h->emb_m = h->emb_tab + d; h->emb_v = h->emb_tab + 2 * d;
    if (h->cfg.adam_mode != SCORE_ADAM_SPARSE) {
        // -1 = no optimizer step has touched the row (its slots are zero, so the dense-gradient semantics leave it where it
        // is): such rows are current by construction; anything that writes a row's slots stamps it with a real step
        CK(cudaMalloc(&h->last_step, sizeof(int32_t) * V));
        CK(cudaMemsetAsync(h->last_step, 0xff, sizeof(int32_t) * V, h->st));
    }
    const int64_t n = h->n_dense;
    CK(cudaMalloc(&h->P, sizeof(float) * n));
    CK(cudaMalloc(&h->G, sizeof(float) * n));
    CK(cudaMalloc(&h->M1, sizeof(float) * n));
    CK(cudaMalloc(&h->V1, sizeof(float) * n));
    CK(cudaMalloc(&h->PG, sizeof(float) * n * kSplits));
    CK(cudaMalloc(&h->flags, n));
    CK(cudaMalloc(&h->Dv, sizeof(float) * (h->n_derived > 0 ? h->n_derived : 4)));
    CK(cudaMemsetAsync(h->Dv, 0, sizeof(float) * (h->n_derived > 0 ? h->n_derived : 4), h->st));
    CK(cudaMemsetAsync(h->P, 0, sizeof(float) * n, h->st));
    CK(cudaMemsetAsync(h->G, 0, sizeof(float) * n, h->st));
    CK(cudaMemsetAsync(h->M1, 0, sizeof(float) * n, h->st));
    CK(cudaMemsetAsync(h->V1, 0, sizeof(float) * n, h->st));
    std::vector<uint8_t> fl(n, 0);
    for (const Tensor& t : h->tensors)
        if (!t.is_emb)
            for (int64_t i = 0; i < t.rows * t.cols; ++i) fl[t.off + i] = t.flags;
    CK(cudaMemcpy(h->flags, fl.data(), n, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->claim_counter, 2 * sizeof(int32_t)));   // ping-pong pair (build_keys / emb_replay)
    CK(cudaMemsetAsync(h->claim_counter, 0, 2 * sizeof(int32_t), h->st));
    CK(cudaMalloc(&h->n_heads_dev, 8 * sizeof(int32_t)));
    CK(cudaMemsetAsync(h->n_heads_dev, 0, 8 * sizeof(int32_t), h->st));
    CK(cudaMalloc(&h->l2sum, sizeof(float) * L2_PARTS));
    CK(cudaMalloc(&h->loss_dev, 2 * sizeof(float)));
    CK(cudaMallocHost(&h->early_host, 4 * sizeof(float)));   // page-locked: the device writes the result packet into it
    memset(h->early_host, 0xff, 4 * sizeof(float));
    CK(cudaMalloc(&h->err_flag, sizeof(int32_t)));
    CK(cudaMemsetAsync(h->err_flag, 0, sizeof(int32_t), h->st));
    CK(cudaMalloc(&h->step_dev, sizeof(ScoreModel::StepParams)));
    h->hyper_dev = &h->step_dev->hp;
    h->bp_dev = &h->step_dev->bp;
    CK(cudaMallocHost(&h->hyper_ring, sizeof(ScoreModel::StepParams) * ScoreModel::kHyperSlots));
    for (int i = 0; i < ScoreModel::kHyperSlots; ++i) CK(cudaEventCreateWithFlags(&h->hyper_ev[i], cudaEventDisableTiming));
    h->hyper_host = &h->hyper_ring[0].hp;
    CK(cudaMallocHost(&h->loss_host, 2 * sizeof(float)));
    CK(cudaMallocHost(&h->err_host, sizeof(int32_t)));
    if (h->cfg.adam_mode == SCORE_ADAM_LAZY) {
        h->alpha_cap = 1 << 20;
        if (const char* e = getenv("SCORE_ALPHA_CAP")) {   // test knob: a short window exercises the re-base
            const long v = atol(e);
            if (v >= 4) h->alpha_cap = v;
        }
        CK(cudaMalloc(&h->alpha_buf, sizeof(float) * h->alpha_cap));
        CK(cudaMemsetAsync(h->alpha_buf, 0, sizeof(float) * h->alpha_cap, h->st));
        h->alpha_hist = h->alpha_buf; h->hist_base = 0;
    }
    return SCORE_OK;
}

// TF-default initial values (score.py:44 truncated_normal; tf.layers.dense glorot_uniform / zeros;
// GRUCell gate bias 1.0; batch_normalization gamma 1, beta 0, moving_mean 0, moving_variance 1)
int init_weights(ScoreModel* h) {
    const uint64_t seed = h->cfg.seed;
    launch_init_trunc_normal(h->st, h->emb, h->cfg.feature_size * h->cfg.eb_dim, seed, 1000u, h->cfg.eb_dim, h->es);
    uint32_t sid = 1;
    for (const Tensor& t : h->tensors) {
        if (t.is_emb) continue;
        float* p = h->P + t.off;
        const int64_t n = t.rows * t.cols;
        const std::string& nm = t.name;
        auto ends = [&](const char* s) { size_t l = strlen(s); return nm.size() >= l && nm.compare(nm.size() - l, l, s) == 0; };
        if (ends("/kernel")) {
            float lim = sqrtf(6.0f / (float)(t.rows + t.cols));
            launch_init_uniform(h->st, p, n, lim, seed, sid++);
        } else if (ends("gates/bias") || nm == "bn1/gamma" || nm == "bn1/moving_variance") {
            launch_fill(h->st, p, n, 1.0f);
        } else {
            launch_fill(h->st, p, n, 0.0f);
        }
    }
    CK(cudaStreamSynchronize(h->st));
    return SCORE_OK;
}

void free_workspace(ScoreModel* h) {
    for (void* p : h->allocs) cudaFree(p);
    h->allocs.clear();
    h->bufs.clear();
    for (auto& kv : h->graphs_train) cudaGraphExecDestroy(kv.second);
    h->graphs_train.clear();
    h->warm_train.clear();
    h->graph_kernels.clear();
    for (auto& kv : h->graphs_begin) cudaGraphExecDestroy(kv.second);
    h->graphs_begin.clear();
    h->warm_begin.clear();
    h->graph_kernels_begin.clear();
    if (h->seg_rows) cudaFree(h->seg_rows);
    if (h->seg_heads) cudaFree(h->seg_heads);
    h->seg_rows = nullptr; h->seg_heads = nullptr; h->seg_cap = 0;
    h->cap_B = 0;
}

template <typename T>
int ws_alloc(ScoreModel* h, T** out, size_t count, const char* name, int dtype = 0) {
    void* p = nullptr;
    size_t bytes = (count ? count : 1) * sizeof(T);
    CK(cudaMalloc(&p, bytes));
    CK(cudaMemsetAsync(p, 0, bytes, h->st));
    h->allocs.push_back(p);
    *out = (T*)p;
    if (name) h->bufs[name] = Buf{p, count, dtype};
    return SCORE_OK;
}

#define WS(ptr, count, name)                                   \
    do {                                                       \
        int rc_ = ws_alloc(h, &(ptr), (size_t)(count), name);  \
        if (rc_) return rc_;                                   \
    } while (0)
#define WSI(ptr, count, name)                                     \
    do {                                                          \
        int rc_ = ws_alloc(h, &(ptr), (size_t)(count), name, 1);  \
        if (rc_) return rc_;                                      \
    } while (0)

int64_t ids_per_sample(const Dims& dm) {
    return (int64_t)dm.T * dm.K * 2 * (dm.fu + dm.fi) + dm.fu + dm.fi;
}

int ensure_workspace(ScoreModel* h, int B) {
    if (B <= h->cap_B) return SCORE_OK;
    CK(cudaStreamSynchronize(h->st));
    free_workspace(h);
    int cap = B;
    const Dims& dm = h->dm;
    const int64_t M = (int64_t)cap * dm.T, N = (int64_t)cap * ids_per_sample(dm);
    const int H = dm.H, K = dm.K, Ds = dm.Ds, Dk = dm.Dk, Dfc = dm.Dfc;
    WSI(h->ids, N, "ids"); WSI(h->label, cap, "label"); WSI(h->length, cap, "length"); WSI(h->keys, N, "keys");
    WSI(h->live, M + 1, "live_slices");
    WS(h->q0, (int64_t)cap * Ds, "q0"); WS(h->c_item, cap, "c_item"); WS(h->c_user, cap, "c_user");
    WS(h->xhg[0], M * dm.ldxs[0], "xhg_user"); WS(h->xhg[1], M * dm.ldxs[1], "xhg_item");
    WS(h->xhc[0], M * dm.ldxs[0], "xhc_user"); WS(h->xhc[1], M * dm.ldxs[1], "xhc_item");
    WS(h->key, M * Dk, "key"); WS(h->save_r, M * 2 * K, "coatt_r"); WS(h->save_w, M * 2 * K, "coatt_w");
    for (int s = 0; s < 2; ++s) {
        const char* sn = s == 0 ? "user" : "item";
        std::string a = std::string("px_") + sn, b = std::string("gru_r_") + sn, c = std::string("gru_u_") + sn,
                    e = std::string("gru_c_") + sn, f = std::string("dpx_") + sn, g = std::string("dx_") + sn;
        WS(h->px[s], M * 3 * H, a.c_str()); WS(h->gr[s], M * H, b.c_str()); WS(h->gu[s], M * H, c.c_str());
        WS(h->gc[s], M * H, e.c_str()); WS(h->dpx[s], M * 3 * H, f.c_str()); WS(h->dx[s], M * dm.Dx[s], g.c_str());
    }
    WS(h->q, (int64_t)cap * Dk, "q"); WS(h->attU, (int64_t)cap * 80, "att_u"); WS(h->qk, M * Dk, "att_qk"); WS(h->f1, M * 80, "att_fc1");
    WS(h->f2, M * 40, "att_fc2"); WS(h->score, M, "score"); WS(h->fc_in, (int64_t)cap * Dfc, "fc_in");
    WS(h->z0, (int64_t)cap * Dfc, "bn1"); WS(h->g1, (int64_t)cap * 200, "fc1"); WS(h->g2, (int64_t)cap * 80, "fc2");
    WS(h->y, cap, "y_pred"); WS(h->loss_b, cap, "loss_b"); WS(h->dlogit, cap, "dlogit");
    WS(h->dg2, (int64_t)cap * 80, "d_fc2"); WS(h->dg1, (int64_t)cap * 200, "d_fc1"); WS(h->dz0, (int64_t)cap * Dfc, "d_bn1");
    WS(h->dfc_in, (int64_t)cap * Dfc, "d_fc_in"); WS(h->ds, M, "d_score"); WS(h->df2, M * 40, "d_att_fc2");
    WS(h->df1, M * 80, "d_att_fc1"); WS(h->sdf1, (int64_t)cap * 80, "d_att_fc1_sum"); WS(h->dqD, (int64_t)cap * Dk, "d_q_keyside");
    WS(h->dkey, M * Dk, "d_key");
    WS(h->dq, (int64_t)cap * Dk, "d_q"); WS(h->dq0, (int64_t)cap * Ds, "d_q0"); WS(h->sdz, M * 2, "sdz");
    WS(h->grad_rows, N * dm.d, "grad_rows");
    h->n_coatt_part = coatt_bwd_num_ctas();
    h->n_target_part = target_bwd_num_ctas();
    WS(h->coatt_part, (int64_t)h->n_coatt_part * (2 * dm.Di + 2 * dm.Du), nullptr);
    WS(h->target_part, (int64_t)h->n_target_part * (dm.Di + dm.Du + 2), nullptr);
    WSI(h->claim_list, 2 * N, nullptr);
    WSI(h->sb.keys[0], N, nullptr); WSI(h->sb.keys[1], N, nullptr);
    WSI(h->sb.vals[0], N, nullptr); WSI(h->sb.vals[1], N, nullptr);
    WSI(h->sb.runs, 8 * N, nullptr);
    WSI(h->sb.runs_long, 4 * emb_runs_long_cap(N), nullptr);
    WS(h->sb.part, emb_runs_part_cap(N) * dm.d, nullptr);
    WSI(h->sb.slotinfo, 4 * emb_runs_part_cap(N), nullptr);
    WSI(h->sb.done, emb_runs_part_cap(N), nullptr);
    {
        uint32_t* hist = nullptr;
        int rc = ws_alloc(h, &hist, sort_hist_elems(N), nullptr);
        if (rc) return rc;
        h->sb.hist = hist;
    }
    // host batches are staged through two buffer sets (upload_batch): set 0 is `ids`
    h->ids_stage[0] = h->ids;
    WSI(h->ids_stage[1], N, nullptr);
    for (int i = 0; i < 2; ++i) { WSI(h->lab_stage[i], cap, nullptr); WSI(h->len_stage[i], cap, nullptr); }
    h->stage_busy[0] = h->stage_busy[1] = false; h->stage_last = -1;
    WSI(h->head_slot, N, nullptr);
    WSI(h->hs_tiles, head_slot_tiles(N) + 1, nullptr);
    h->cap_B = cap;
    CK(cudaStreamSynchronize(h->st));
    return SCORE_OK;
}

void set_batch_dims(ScoreModel* h, int B) {
    Dims& dm = h->dm;
    dm.B = B;
    const int64_t M = (int64_t)B * dm.T;
    dm.nrows = dm.K * (2 * dm.fi + 2 * dm.fu);
    dm.off_tu = M * dm.nrows;
    dm.off_ti = dm.off_tu + (int64_t)B * dm.fu;
    dm.N = dm.off_ti + (int64_t)B * dm.fi;
}

// Make the batch visible to the device.  A device batch is consumed where it lies (pointer table only); a host batch
// is copied into the handle's staging buffers first.  The table reaches the device with the next upload_hyper().
int upload_batch(ScoreModel* h, const ScoreBatch* b) {
    const Dims& dm = h->dm;
    const int B = b->batch_size;
    const int64_t M = (int64_t)B * dm.T;
    BatchPtrs& bp = h->bp_cur;
    if (b->on_device) {
        h->stage_last = -1;
        bp.u1 = b->user_1hop; bp.u2 = b->user_2hop; bp.i1 = b->item_1hop; bp.i2 = b->item_2hop;
        bp.tu = b->target_user; bp.ti = b->target_item; bp.label = b->label; bp.length = b->length;
        return SCORE_OK;
    }
    // Host batch: copied into one of two staging sets on the copy stream, so the copy of the NEXT batch can run while the
    // previous step is still in its backward pass (its build_keys - the only reader of a staging set - ran long ago; the
    // event guards the asynchronous callers that run several steps ahead).
    const size_t s_i = sizeof(int32_t);
    const int64_t n_i = M * dm.K * dm.fi, n_u = M * dm.K * dm.fu;
    const int sidx = h->stage_idx ^= 1;
    cudaStream_t cs = h->st_h2d;
    if (h->stage_busy[sidx]) CK(cudaStreamWaitEvent(cs, h->ev_stage_free[sidx], 0));
    int32_t* st_u1 = h->ids_stage[sidx]; int32_t* st_u2 = st_u1 + n_i; int32_t* st_i1 = st_u2 + n_u; int32_t* st_i2 = st_i1 + n_u;
    int32_t* st_tu = st_i2 + n_i; int32_t* st_ti = st_tu + (int64_t)B * dm.fu;
    CK(cudaMemcpyAsync(st_u1, b->user_1hop, s_i * n_i, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(st_u2, b->user_2hop, s_i * n_u, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(st_i1, b->item_1hop, s_i * n_u, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(st_i2, b->item_2hop, s_i * n_i, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(st_tu, b->target_user, s_i * B * dm.fu, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(st_ti, b->target_item, s_i * B * dm.fi, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(h->lab_stage[sidx], b->label, s_i * B, cudaMemcpyHostToDevice, cs));
    CK(cudaMemcpyAsync(h->len_stage[sidx], b->length, s_i * B, cudaMemcpyHostToDevice, cs));
    CK(cudaEventRecord(h->ev_h2d[sidx], cs));
    CK(cudaStreamWaitEvent(h->st, h->ev_h2d[sidx], 0));
    h->stage_last = sidx;
    bp.u1 = st_u1; bp.u2 = st_u2; bp.i1 = st_i1; bp.i2 = st_i2; bp.tu = st_tu; bp.ti = st_ti;
    bp.label = h->lab_stage[sidx]; bp.length = h->len_stage[sidx];
    return SCORE_OK;
}

// after the kernels that read the current batch have been enqueued: the staging set may be refilled behind them
void stage_release(ScoreModel* h) {
    if (h->stage_last < 0) return;
    cudaEventRecord(h->ev_stage_free[h->stage_last], h->st);
    h->stage_busy[h->stage_last] = true;
    h->stage_last = -1;
}

// ------------------------------------------------------------------------------------------ GEMM helpers
// dW[Kl, Nl] (+ db[Nl]) = A[M, Kl]^T dC[M, Nl], into the kSplits partial planes of PG
GemmArgs gemm_bwd_weight_args(ScoreModel* h, const float* A, int lda, const float* dC, int lddc, int64_t w_off,
                              int64_t b_off, int M, int Kl, int Nl) {
    GemmArgs g{};
    g.A = A; g.a_rs = 1; g.a_cs = lda;
    g.B = dC; g.b_rs = lddc; g.b_cs = 1;
    g.C = h->PG + w_off; g.c_rs = Nl; g.M = Kl; g.N = Nl; g.K = M; g.epi = EPI_SPLIT;
    g.splits = kSplits; g.c_split_stride = h->n_dense;
    g.colsum = (b_off >= 0) ? h->PG + b_off : nullptr; g.colsum_split_stride = h->n_dense;
    g.hp = h->hyper_dev;
    return g;
}
// several weight-gradient problems in one side-stream launch: weight gradients are off the critical path (they only
// feed the final reduce), ordered after the producer of dC by an event
void gemm_bwd_weight_batch(ScoreModel* h, const GemmArgs* list, int n) {
    cudaEvent_t e = h->ev_pool[h->ev_next++ & 15];
    cudaEventRecord(e, h->st);
    cudaStreamWaitEvent(h->st_w, e, 0);
    launch_gemm_batch(h->st_w, list, n);
}
void probe_begin(ScoreModel* h, int p, cudaStream_t s) {
    if (h->probes_on) cudaEventRecordWithFlags(h->pr_beg[p], s, h->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
}
void probe_end(ScoreModel* h, int p, cudaStream_t s) {
    if (h->probes_on) cudaEventRecordWithFlags(h->pr_end[p], s, h->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
}

enum StepMode { MODE_TRAIN = 0, MODE_EVAL = 1, MODE_FWDBWD = 2, MODE_BEGIN = 3 };

int key_bits(int64_t V);
// radix sort of the batch's (key, position) pairs + run descriptors on the sort stream, forked at `after`.
// part 0: everything; part 1: the first pass only (before the gather kernel); part 2: the remaining passes + descriptors
void enqueue_sort_branch(ScoreModel* h, cudaEvent_t after, int part = 0) {
    const Dims& dm = h->dm;
    const int np = sort_num_passes(key_bits(dm.V));
    cudaStreamWaitEvent(h->st2, after, 0);
    if (part != 2) probe_begin(h, PR_SORT, h->st2);
    if (part == 1) { h->sort_out = launch_sort_passes(h->st2, h->sb, h->keys, dm.N, 0, 1); return; }
    h->sort_out = launch_sort_passes(h->st2, h->sb, h->keys, dm.N, part == 2 ? 1 : 0, np);
    launch_emb_runs(h->st2, h->sb.keys[h->sort_out], h->sb.vals[h->sort_out], dm.N, h->sb.runs, h->sb.runs_long, h->n_heads_dev, h->sb.slotinfo);
    probe_end(h, PR_SORT, h->st2);
    if (h->sort_dp) {
        // data-parallel half-step: the host sizes the exchange from the unique-row count - publish it as soon as it
        // exists (score_dp_local_count) - and the export needs every run's rank among the runs (compact sorted list)
        launch_dp_count(h->st2, h->n_heads_dev, h->hyper_dev, h->cnt_slot);
        cudaEventRecordWithFlags(h->ev_counts, h->st2, h->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
        launch_head_slots(h->st2, h->sb.keys[h->sort_out], dm.N, h->hs_tiles, h->head_slot);
    }
    cudaEventRecord(h->ev_join, h->st2);
}

// forward graph of class SCORE (score.py:188-224); everything enqueued on h->st
void enqueue_forward(ScoreModel* h, bool will_bwd) {
    const Dims& dm = h->dm;
    const int B = dm.B, T = dm.T, H = dm.H, Ds = dm.Ds, Dk = dm.Dk, Dfc = dm.Dfc;
    const int M = B * T;
    Names nm = role_names(dm.model_type);
    auto W = [&](const std::string& n) { return pp(h, (n + "/kernel").c_str()); };
    auto Bi = [&](const std::string& n) { return pp(h, (n + "/bias").c_str()); };

    // side stream: derived weights of the fused chains, then 0.5*sum(v^2) (only feeds the loss scalar)
    cudaStreamWaitEvent(h->st_w, h->ev_fork, 0);
    if (will_bwd) cudaMemsetAsync(h->PG, 0, sizeof(float) * h->n_dense * kSplits, h->st_w);   // weight-gradient partial planes
    launch_prep_weights(h->st_w, h->prep, h->P, h->Dv);
    cudaEventRecord(h->ev_prep, h->st_w);
    launch_l2_sum(h->st_w, h->P, h->flags, (int)h->n_dense, h->l2sum);
    cudaEventRecord(h->ev_l2, h->st_w);

    const int mt = dm.model_type;
    // RCA sums the neighbors (score.py:266-269); RIA feeds the final GRU states to the MLP (score.py:244-249);
    // RRN does both on the 1-hop tensors only (slice_model.py:158-168)
    const bool has_coatt = mt != SCORE_MODEL_RCA && mt != SCORE_MODEL_RRN;
    const bool has_att = mt != SCORE_MODEL_RIA && mt != SCORE_MODEL_RRN;
    TargetArgs ta{};
    ta.emb = h->emb_fwd; ta.es = h->es_fwd; ta.keys = h->keys_fwd;
    if (has_coatt) { ta.w_item = W(nm.co_item); ta.b_item = Bi(nm.co_item); ta.w_user = W(nm.co_user); ta.b_user = Bi(nm.co_user); }
    ta.q0 = h->q0; ta.fc_in = h->fc_in; ta.fc_off = Dfc - Ds; ta.c_item = h->c_item; ta.c_user = h->c_user;
    launch_target_fwd(h->st, dm, ta);
    if (has_att) {   // query side of the attention (per sample): off the critical path, next to the gather and the GRUs
        cudaEventRecord(h->ev_tgt, h->st);
        cudaStreamWaitEvent(h->st_w, h->ev_tgt, 0);
        AttQArgs qa{};
        qa.B = B; qa.Ds = Ds; qa.Dk = Dk; qa.q0 = h->q0; qa.wq = W(nm.att_q); qa.bq = Bi(nm.att_q);
        qa.Wac = h->Dv + h->dv_Wac; qa.b1 = Bi(nm.att1); qa.q = h->q; qa.U = h->attU;
        launch_att_q(h->st_w, qa);
        cudaEventRecord(h->ev_q, h->st_w);
    }

    CoattArgs ca{};
    ca.emb = h->emb_fwd; ca.es = h->es_fwd; ca.keys = h->keys_fwd; ca.length = h->length;
    if (has_coatt) { ca.w_item = W(nm.co_item); ca.w_user = W(nm.co_user); }
    ca.sum_pool = has_coatt ? 0 : 1;
    ca.live = dm.T <= 255 ? h->live : nullptr;
    ca.c_item = h->c_item; ca.c_user = h->c_user;
    ca.xhg_u = h->xhg[0]; ca.xhc_u = h->xhc[0]; ca.xhg_i = h->xhg[1]; ca.xhc_i = h->xhc[1];
    ca.key = h->key; ca.ldkey = Dk; ca.key_off = 2 * H; ca.save_r = h->save_r; ca.save_w = h->save_w;
    probe_begin(h, PR_COATT_FWD, h->st);
    launch_coatt_fwd(h->st, dm, ca);
    probe_end(h, PR_COATT_FWD, h->st);
    if (h->sort_deferred) {   // remaining passes of the sort: after the gather (enqueue_step)
        cudaEventRecord(h->ev_keys, h->st);
        enqueue_sort_branch(h, h->ev_keys, 2);
        h->sort_deferred = false;
    }

    probe_begin(h, PR_FWD_DENSE, h->st);
    const char* sides[2] = {"gru_user_side", "gru_item_side"};
    GruArgs ga{};
    ga.length = h->length;
    RowGemmArgs pxl[2];
    for (int s = 0; s < 2; ++s) {
        std::string g = std::string(sides[s]) + "/gru_cell/gates", c = std::string(sides[s]) + "/gru_cell/candidate";
        // input part of both matmuls for all (b,t) rows at once: px = x [Wg_x | Wc_x]  (both sides, one launch)
        pxl[s] = RowGemmArgs{h->xhg[s], dm.ldxs[s], h->Dv + h->dv_Wx[s], 3 * H, nullptr, h->px[s], 3 * H, M, dm.Dx[s], 3 * H};
        ga.px[s] = h->px[s]; ga.wg[s] = W(g); ga.bg[s] = Bi(g); ga.wc[s] = W(c); ga.bc[s] = Bi(c);
        ga.xhg[s] = h->xhg[s]; ga.xhc[s] = h->xhc[s]; ga.r[s] = h->gr[s]; ga.u[s] = h->gu[s]; ga.c[s] = h->gc[s];
    }
    cudaStreamWaitEvent(h->st, h->ev_prep, 0);   // the derived weights are rebuilt on the side stream
    launch_rowgemm(h->st, pxl, 2);
    ga.out = h->key; ga.ldout = Dk;
    ga.last = has_att ? nullptr : h->fc_in; ga.ldlast = has_att ? 0 : Dfc;   // RIA: fc_in = [user_last | item_last | ...]
    launch_gru_fwd(h->st, dm, ga);

    // attention over the T slices + attentive pooling (score.py:169-186, 214-216): one fused row-tile chain
    if (has_att) {
    cudaStreamWaitEvent(h->st, h->ev_q, 0);
    AttFwd2Args fa2{};
    fa2.B = B; fa2.T = T; fa2.Dk = Dk; fa2.H = H; fa2.length = h->length; fa2.q = h->q; fa2.U = h->attU; fa2.key = h->key;
    fa2.W1e = h->Dv + h->dv_W1e; fa2.w2 = W(nm.att2); fa2.b2 = Bi(nm.att2); fa2.w3 = W(nm.att3); fa2.b3 = Bi(nm.att3);
    fa2.qk = h->qk; fa2.f1 = h->f1; fa2.f2 = h->f2; fa2.score = h->score; fa2.fc_in = h->fc_in; fa2.ldfc = Dfc;
    fa2.model_type = dm.model_type;
    launch_att_fwd2(h->st, fa2);
    }

    // build_fc_net + log-loss (score.py:68-81): one fused row-tile chain
    FcArgs fa{};
    fa.B = B; fa.F = Dfc; fa.fc_in = h->fc_in;
    fa.gamma = pp(h, "bn1/gamma"); fa.beta = pp(h, "bn1/beta"); fa.mean = pp(h, "bn1/moving_mean"); fa.var = pp(h, "bn1/moving_variance");
    fa.w1 = pp(h, "fc1/kernel"); fa.b1 = pp(h, "fc1/bias"); fa.w2 = pp(h, "fc2/kernel"); fa.b2 = pp(h, "fc2/bias");
    fa.w3 = pp(h, "fc3/kernel"); fa.b3 = pp(h, "fc3/bias"); fa.label = h->label; fa.hp = h->hyper_dev;
    fa.z0 = h->z0; fa.g1 = h->g1; fa.g2 = h->g2; fa.y = h->y; fa.loss_b = h->loss_b; fa.dlogit = h->dlogit;
    launch_fc_fwd(h->st, fa);
    probe_end(h, PR_FWD_DENSE, h->st);
    // the loss scalar only leaves the device at the end of the step: reduce it on the side stream (behind l2_sum, same
    // stream) instead of between the forward and the backward chain
    cudaEventRecord(h->ev_fc, h->st);
    cudaStreamWaitEvent(h->st_w, h->ev_fc, 0);
    launch_loss_final(h->st_w, B, h->loss_b, h->l2sum, h->hyper_dev, h->loss_dev, h->err_flag, h->early_on ? h->early_host : nullptr);
    if (!will_bwd) {   // no backward: nothing else joins the side stream
        cudaEventRecord(h->ev_w, h->st_w);
        cudaStreamWaitEvent(h->st, h->ev_w, 0);
    }
}

// fused_adam: the dense Adam step rides on the final gradient reduce (single-GPU training step)
// scheduling knobs (measured on B200, profiles/README.md); SCORE_SCHED=0 restores the serial order
static int sched_flags() {
    static int f = -1;
    if (f < 0) { const char* e = getenv("SCORE_SCHED"); f = e ? atoi(e) : 15; }
    return f;
}
static bool side_qb() { return (sched_flags() & 1) != 0; }        // att_qb on the side stream
static bool side_reduce() { return (sched_flags() & 2) != 0; }    // final dense reduce + Adam next to the embedding update
static bool sort_after_gather() { return (sched_flags() & 4) != 0; }   // sort branch forks after the gather kernel
static bool early_return() { return (sched_flags() & 8) != 0; }        // synchronous train returns when the loss is ready

void enqueue_backward(ScoreModel* h, bool fused_adam = false, bool defer_join = false) {
    DenseAdamArgs adam_args{h->P, h->M1, h->V1, h->flags, h->hyper_dev, h->alpha_hist};
    const DenseAdamArgs* adam = fused_adam ? &adam_args : nullptr;
    const Dims& dm = h->dm;
    const int B = dm.B, T = dm.T, H = dm.H, Ds = dm.Ds, Dk = dm.Dk, Dfc = dm.Dfc;
    const int M = B * T;
    Names nm = role_names(dm.model_type);
    auto W = [&](const std::string& n) { return pp(h, (n + "/kernel").c_str()); };
    auto Wo = [&](const std::string& n) { return po(h, (n + "/kernel").c_str()); };
    auto Bo = [&](const std::string& n) { return po(h, (n + "/bias").c_str()); };

    probe_begin(h, PR_BWD_DENSE, h->st);

    // prediction MLP: fused backward data chain on the main stream, weight gradients on the side stream
    FcBwdArgs fb{};
    fb.B = B; fb.F = Dfc; fb.dlogit = h->dlogit; fb.g2 = h->g2; fb.g1 = h->g1;
    fb.w3 = W("fc3"); fb.w2 = W("fc2"); fb.w1 = W("fc1"); fb.gamma = pp(h, "bn1/gamma"); fb.var = pp(h, "bn1/moving_variance");
    fb.w2t = h->Dv + h->dv_fc2T; fb.w1t = h->Dv + h->dv_fc1T;
    fb.hp = h->hyper_dev; fb.dg2 = h->dg2; fb.dg1 = h->dg1; fb.dz0 = h->dz0; fb.dfc_in = h->dfc_in;
    launch_fc_bwd(h->st, fb);
    {   // weight gradients go to the side stream in groups: one launch per group of problems that become ready together
        GemmArgs l[3];
        l[0] = gemm_bwd_weight_args(h, h->z0, Dfc, h->dg1, 200, Wo("fc1"), Bo("fc1"), B, Dfc, 200);
        l[1] = gemm_bwd_weight_args(h, h->g1, 200, h->dg2, 80, Wo("fc2"), Bo("fc2"), B, 200, 80);
        l[2] = gemm_bwd_weight_args(h, h->g2, 80, h->dlogit, 1, Wo("fc3"), Bo("fc3"), B, 80, 1);
        gemm_bwd_weight_batch(h, l, 3);
    }
    launch_bn_param_grads(h->st_w, B, Dfc, h->fc_in, h->dz0, pp(h, "bn1/moving_mean"), pp(h, "bn1/moving_variance"),
                          h->PG + po(h, "bn1/gamma"), h->PG + po(h, "bn1/beta"), kSplits, h->n_dense);

    // attention: pooling + softmax + MLP backward in one fused chain, then the per-sample query side
    const int mt = dm.model_type;
    const bool has_coatt = mt != SCORE_MODEL_RCA && mt != SCORE_MODEL_RRN, has_att = mt != SCORE_MODEL_RIA && mt != SCORE_MODEL_RRN;
    const int64_t blk = (int64_t)Dk * 80;   // dense_3/kernel row blocks: Wa | Wb | Wc | Wd
    if (has_att) {
        AttBwd2Args ab{};
        ab.B = B; ab.T = T; ab.Dk = Dk; ab.H = H; ab.length = h->length; ab.q = h->q; ab.key = h->key; ab.f1 = h->f1;
        ab.f2 = h->f2; ab.score = h->score; ab.dfc_in = h->dfc_in; ab.ldfc = Dfc; ab.model_type = dm.model_type;
        ab.w3 = W(nm.att3); ab.W2T = h->Dv + h->dv_W2T; ab.W1eT = h->Dv + h->dv_W1eT;
        ab.ds = h->ds; ab.df2 = h->df2; ab.df1 = h->df1; ab.dkey = h->dkey; ab.sdf1 = h->sdf1; ab.dqD = h->dqD;
        launch_att_bwd2(h->st, ab);
        GemmArgs w1l[4];
        w1l[0] = gemm_bwd_weight_args(h, h->key, Dk, h->df1, 80, Wo(nm.att1) + blk, -1, M, Dk, 80);      // dWb = key^T df1
        w1l[1] = gemm_bwd_weight_args(h, h->qk, Dk, h->df1, 80, Wo(nm.att1) + 3 * blk, -1, M, Dk, 80);   // dWd = (q*key)^T df1
        w1l[2] = gemm_bwd_weight_args(h, h->f1, 80, h->df2, 40, Wo(nm.att2), Bo(nm.att2), M, 80, 40);
        w1l[3] = gemm_bwd_weight_args(h, h->f2, 40, h->ds, 1, Wo(nm.att3), Bo(nm.att3), M, 40, 1);
        // query side of the attention backward (per sample): only target_bwd and two weight gradients consume it, so it
        // leaves the critical path (the GRU backward does not wait for it) and runs first on the side stream
        AttQbArgs qb{};
        qb.B = B; qb.Ds = Ds; qb.Dk = Dk; qb.sdf1 = h->sdf1; qb.dqD = h->dqD; qb.WacT = h->Dv + h->dv_WacT;
        qb.WqT = h->Dv + h->dv_WqT; qb.dq = h->dq; qb.dq0 = h->dq0;
        if (side_qb()) {
            cudaEventRecord(h->ev_att, h->st);
            cudaStreamWaitEvent(h->st_w, h->ev_att, 0);
            launch_att_qb(h->st_w, qb);
            cudaEventRecord(h->ev_qb, h->st_w);
        }
        gemm_bwd_weight_batch(h, w1l, 4);
        if (!side_qb()) launch_att_qb(h->st, qb);
        GemmArgs wql[2];
        wql[0] = gemm_bwd_weight_args(h, h->q, Dk, h->sdf1, 80, Wo(nm.att1), Bo(nm.att1), B, Dk, 80);    // dWa = q^T sum_t df1
        wql[1] = gemm_bwd_weight_args(h, h->q0, Ds, h->dq, Dk, Wo(nm.att_q), Bo(nm.att_q), B, Ds, Dk);
        gemm_bwd_weight_batch(h, wql, 2);
    }

    // GRUs
    const char* sides[2] = {"gru_user_side", "gru_item_side"};
    GruBwdArgs gb{};
    gb.length = h->length;
    gb.dout = has_att ? h->dkey : nullptr; gb.lddout = Dk;                       // RIA: only the final states are consumed
    gb.dlast = has_att ? nullptr : h->dfc_in; gb.lddlast = has_att ? 0 : Dfc;
    for (int s = 0; s < 2; ++s) {
        std::string g = std::string(sides[s]) + "/gru_cell/gates", c = std::string(sides[s]) + "/gru_cell/candidate";
        gb.wg[s] = W(g); gb.wc[s] = W(c); gb.xhg[s] = h->xhg[s];
        gb.r[s] = h->gr[s]; gb.u[s] = h->gu[s]; gb.c[s] = h->gc[s]; gb.dpx[s] = h->dpx[s];
    }
    launch_gru_bwd(h->st, dm, gb);
    GemmArgs wl[4];
    RowGemmArgs dxl[2];
    for (int s = 0; s < 2; ++s) {
        std::string g = std::string(sides[s]) + "/gru_cell/gates", c = std::string(sides[s]) + "/gru_cell/candidate";
        wl[2 * s] = gemm_bwd_weight_args(h, h->xhg[s], dm.ldxs[s], h->dpx[s], 3 * H, Wo(g), Bo(g), M, dm.ldxs[s], 2 * H);
        wl[2 * s + 1] = gemm_bwd_weight_args(h, h->xhc[s], dm.ldxs[s], h->dpx[s] + 2 * H, 3 * H, Wo(c), Bo(c), M, dm.ldxs[s], H);
        // dx = dpx [Wg_x | Wc_x]^T
        dxl[s] = RowGemmArgs{h->dpx[s], 3 * H, h->Dv + h->dv_WxT[s], dm.Dx[s], nullptr, h->dx[s], dm.Dx[s], M, 3 * H, dm.Dx[s]};
    }
    gemm_bwd_weight_batch(h, wl, 4);
    launch_rowgemm(h->st, dxl, 2);
    probe_end(h, PR_BWD_DENSE, h->st);
    // co-attention + gather backward: per-position embedding gradient rows
    CoattBwdArgs cb{};
    cb.emb = h->emb_fwd; cb.es = h->es_fwd; cb.keys = h->keys_fwd; cb.length = h->length;
    if (has_coatt) { cb.w_item = W(nm.co_item); cb.w_user = W(nm.co_user); }
    cb.sum_pool = has_coatt ? 0 : 1;
    cb.live = dm.T <= 255 ? h->live : nullptr;
    cb.save_r = h->save_r; cb.save_w = h->save_w;
    cb.dxu = h->dx[0]; cb.dxi = h->dx[1];
    cb.dkey = (has_att && has_coatt) ? h->dkey : nullptr;   // RIA never consumes atten_info; RCA has none
    cb.ldkey = Dk; cb.key_off = 2 * H;
    cb.grad_rows = h->grad_rows; cb.sdz = h->sdz; cb.partials = h->coatt_part; cb.n_partials = h->n_coatt_part;
    probe_begin(h, PR_COATT_BWD, h->st);
    launch_coatt_bwd(h->st, dm, cb);
    probe_end(h, PR_COATT_BWD, h->st);
    TargetBwdArgs tb{};
    tb.length = h->length;
    if (has_coatt) { tb.w_item = W(nm.co_item); tb.w_user = W(nm.co_user); }
    tb.q0 = h->q0; tb.dq0 = has_att ? h->dq0 : nullptr;
    tb.dfc_in = h->dfc_in; tb.fc_off = Dfc - Ds; tb.ldfc = Dfc; tb.sdz = h->sdz; tb.grad_rows = h->grad_rows;
    tb.partials = h->target_part; tb.n_partials = h->n_target_part;
    if (has_att && side_qb()) cudaStreamWaitEvent(h->st, h->ev_qb, 0);
    launch_target_bwd(h->st, dm, tb);
    // Final reduce of the dense gradients (+ fused dense Adam).  defer_join: it runs on the side stream, next to the
    // embedding update that follows on the main stream (neither reads what the other writes); the caller joins with
    // join_dense() afterwards.
    cudaStream_t rs = h->st;
    if (defer_join) {
        cudaEventRecord(h->ev_att, h->st);
        cudaStreamWaitEvent(h->st_w, h->ev_att, 0);
        rs = h->st_w;
    }
    if (has_coatt)
        launch_coatt_grad_reduce(rs, dm, h->coatt_part, h->n_coatt_part, h->target_part, h->n_target_part,
                                 h->PG + Wo(nm.co_item), h->PG + Bo(nm.co_item), h->PG + Wo(nm.co_user),
                                 h->PG + Bo(nm.co_user));
    if (!defer_join) {
        cudaEventRecord(h->ev_w, h->st_w);
        cudaStreamWaitEvent(h->st, h->ev_w, 0);
    }
    if (has_att) {   // dWc = dWa - dWb (attn.cu header)
        const int64_t w1 = po(h, (nm.att1 + "/kernel").c_str()), blk = (int64_t)Dk * 80;
        launch_reduce_partials(rs, h->PG, kSplits, (int)h->n_dense, h->G, w1 + 2 * blk, w1, w1 + blk, (int)blk, adam);
    } else {
        launch_reduce_partials(rs, h->PG, kSplits, (int)h->n_dense, h->G, -1, 0, 0, 0, adam);
    }
    if (defer_join) cudaEventRecord(h->ev_w, h->st_w);
}
void join_dense(ScoreModel* h) { cudaStreamWaitEvent(h->st, h->ev_w, 0); }

int key_bits(int64_t V) {
    int bits = 1;
    while (((int64_t)1 << bits) < V) ++bits;
    return bits;
}

int ensure_seg(ScoreModel* h, int64_t N) {
    if (N <= h->seg_cap) return SCORE_OK;
    if (h->seg_rows) { cudaFree(h->seg_rows); cudaFree(h->seg_heads); }
    CK(cudaMalloc(&h->seg_rows, sizeof(float) * N * h->dm.d));
    CK(cudaMalloc(&h->seg_heads, sizeof(int32_t) * N));
    h->seg_cap = N;
    return SCORE_OK;
}

// Everything of one step that runs on the device, in stream order (capturable).
void enqueue_step(ScoreModel* h, int mode) {
    const Dims& dm = h->dm;
    const bool train = (mode == MODE_TRAIN);
    const bool need_bwd = (mode != MODE_EVAL);
    h->emb_fwd = h->emb; h->es_fwd = h->es; h->keys_fwd = h->keys;
    probe_begin(h, PR_STEP, h->st);
    // The sort depends on ids only and runs on the side stream under forward/backward.  With sort_after_gather() it
    // pauses while the gather kernel - the one bandwidth-bound kernel of the forward pass - has the SMs to itself: the
    // first pass runs next to the (instruction-bound) lazy replay, the other passes are forked after the gather.
    h->sort_deferred = need_bwd && sort_after_gather();
    h->sort_dp = false;
    h->early_on = train;
    if (h->cfg.adam_mode == SCORE_ADAM_LAZY) {
        probe_begin(h, PR_CATCHUP, h->st);
        ClaimArgs ca{h->last_step, h->hyper_dev, h->claim_list, h->claim_counter, 1};
        launch_build_keys(h->st, dm, h->bp_dev, h->keys, h->label, h->length, h->err_flag, &ca, h->live);
        if (h->sort_deferred) { cudaEventRecord(h->ev_keys, h->st); enqueue_sort_branch(h, h->ev_keys, 1); }
        launch_emb_replay(h->st, h->claim_list, h->claim_counter, dm.N, h->emb, h->emb_m, h->emb_v, dm.d, h->es, h->alpha_hist,
                          h->hyper_dev, 1);
        probe_end(h, PR_CATCHUP, h->st);
    } else {
        launch_build_keys(h->st, dm, h->bp_dev, h->keys, h->label, h->length, h->err_flag, nullptr, h->live);
        if (h->sort_deferred) { cudaEventRecord(h->ev_keys, h->st); enqueue_sort_branch(h, h->ev_keys, 1); }
    }
    cudaEventRecord(h->ev_fork, h->st);
    if (need_bwd && !h->sort_deferred) enqueue_sort_branch(h, h->ev_fork);
    enqueue_forward(h, need_bwd);
    const bool defer = train && side_reduce();
    if (need_bwd) {
        enqueue_backward(h, train, defer);
        cudaStreamWaitEvent(h->st, h->ev_join, 0);
    }
    if (mode == MODE_FWDBWD) {
        launch_add_l2(h->st, h->G, h->P, h->flags, (int)h->n_dense, h->hyper_dev);
        cudaMemsetAsync(h->seg_heads, 0, sizeof(int32_t) * dm.N, h->st);
        EmbUpdateArgs ea{};
        ea.skeys = h->sb.keys[h->sort_out]; ea.spos = h->sb.vals[h->sort_out]; ea.n = dm.N;
        ea.runs = h->sb.runs; ea.runs_long = h->sb.runs_long; ea.long_cap = emb_runs_long_cap(dm.N); ea.counters = h->n_heads_dev;
        ea.part = h->sb.part; ea.slotinfo = h->sb.slotinfo; ea.done = h->sb.done;
        ea.grad_rows = h->grad_rows; ea.d = dm.d; ea.hp = h->hyper_dev; ea.mode = 1;
        ea.out_rows = h->seg_rows; ea.out_heads = h->seg_heads;
        launch_emb_update(h->st, ea);
    }
    if (train) {   // the dense Adam step was fused into the gradient reduce (enqueue_backward)
        EmbUpdateArgs ea{};
        ea.skeys = h->sb.keys[h->sort_out]; ea.spos = h->sb.vals[h->sort_out]; ea.n = dm.N;
        ea.runs = h->sb.runs; ea.runs_long = h->sb.runs_long; ea.long_cap = emb_runs_long_cap(dm.N); ea.counters = h->n_heads_dev;
        ea.part = h->sb.part; ea.slotinfo = h->sb.slotinfo; ea.done = h->sb.done;
        ea.grad_rows = h->grad_rows; ea.d = dm.d;
        ea.emb = h->emb; ea.m = h->emb_m; ea.v = h->emb_v; ea.es = h->es; ea.last_step = h->last_step;
        ea.alpha_hist = h->alpha_hist;
        ea.hp = h->hyper_dev; ea.mode = 0;
        h->last_sorted = ea.skeys; h->last_sorted_n = ea.n;
        probe_begin(h, PR_EMB_UPDATE, h->st);
        launch_emb_update(h->st, ea);
        probe_end(h, PR_EMB_UPDATE, h->st);
        if (h->cfg.adam_mode == SCORE_ADAM_DENSE)
            launch_emb_dense_sweep(h->st, h->emb, h->emb_m, h->emb_v, h->last_step, dm.V, dm.d, h->es, h->hyper_dev);
    }
    if (defer) join_dense(h);
    probe_end(h, PR_STEP, h->st);
}

int check_batch(ScoreModel* h, const ScoreBatch* b) {
    if (!b) return fail(h, SCORE_ERR_ARG, "batch is NULL");
    if (b->batch_size <= 0) return fail(h, SCORE_ERR_ARG, "batch_size must be positive");
    if (!b->user_1hop || !b->user_2hop || !b->item_1hop || !b->item_2hop || !b->target_user || !b->target_item ||
        !b->label || !b->length)
        return fail(h, SCORE_ERR_ARG, "batch has a NULL id array");
    if ((int64_t)b->batch_size * ids_per_sample(h->dm) >= ((int64_t)1 << 31))
        return fail(h, SCORE_ERR_ARG, "batch too large: position index must fit int32");
    return SCORE_OK;
}

void fill_hyper(ScoreModel* h, int B, float lr, float reg_lambda, float keep_prob, int train, int global_batch) {
    Hyper& hp = *h->hyper_host;
    hp.lr = lr; hp.reg_lambda = reg_lambda; hp.keep_prob = keep_prob;
    const float one = 1.0f;
    hp.alpha = lr * sqrtf(one - h->beta2_power) / (one - h->beta1_power);
    hp.inv_batch = 1.0f / (float)(global_batch > 0 ? global_batch : B);
    hp.seed_lo = (uint32_t)h->cfg.seed; hp.seed_hi = (uint32_t)(h->cfg.seed >> 32);
    hp.step = h->step + 1; hp.batch = B; hp.train = train; hp.seq = h->hyper_seq++; hp.sample_base = h->sample_base;
}

// fill the next ring slot and enqueue its H2D copy
int upload_hyper(ScoreModel* h, int B, float lr, float reg_lambda, float keep_prob, int train, int global_batch) {
    const int slot = h->hyper_next++ % ScoreModel::kHyperSlots;
    if (h->hyper_used[slot]) CK(cudaEventSynchronize(h->hyper_ev[slot]));
    h->hyper_host = &h->hyper_ring[slot].hp;
    fill_hyper(h, B, lr, reg_lambda, keep_prob, train, global_batch);
    h->hyper_ring[slot].bp = h->bp_cur;
    CK(cudaMemcpyAsync(h->step_dev, h->hyper_ring + slot, sizeof(ScoreModel::StepParams), cudaMemcpyHostToDevice, h->st));
    CK(cudaEventRecord(h->hyper_ev[slot], h->st));
    h->hyper_used[slot] = true;
    return SCORE_OK;
}

// new batch pointer table, SAME hyper-parameters: a begun step whose optimizer half is still to be enqueued keeps
// reading its own alpha / step / reg_lambda (row-sharded pipelining: the next batch is planned before the finish)
int upload_bp_keep_hyper(ScoreModel* h) {
    const Hyper keep = *h->hyper_host;
    const int slot = h->hyper_next++ % ScoreModel::kHyperSlots;
    if (h->hyper_used[slot]) CK(cudaEventSynchronize(h->hyper_ev[slot]));
    h->hyper_ring[slot].hp = keep;
    h->hyper_host = &h->hyper_ring[slot].hp;
    h->hyper_ring[slot].bp = h->bp_cur;
    CK(cudaMemcpyAsync(h->step_dev, h->hyper_ring + slot, sizeof(ScoreModel::StepParams), cudaMemcpyHostToDevice, h->st));
    CK(cudaEventRecord(h->hyper_ev[slot], h->st));
    h->hyper_used[slot] = true;
    return SCORE_OK;
}

int flush_lazy(ScoreModel* h) {
    if (h->cfg.adam_mode == SCORE_ADAM_LAZY && h->step > 0) {
        launch_emb_catchup_all(h->st, h->emb, h->emb_m, h->emb_v, h->last_step, h->dm.V, h->dm.d, h->es, h->alpha_hist, h->step);
        CK(cudaStreamSynchronize(h->st));
    }
    return SCORE_OK;
}

void drop_graphs(ScoreModel* h) {
    for (auto& kv : h->graphs_train) cudaGraphExecDestroy(kv.second);
    h->graphs_train.clear(); h->graph_kernels.clear();
    for (auto& kv : h->graphs_begin) cudaGraphExecDestroy(kv.second);
    h->graphs_begin.clear(); h->graph_kernels_begin.clear();
}

// LAZY Adam: declare every row current at h->step and restart the alpha window there.  Callers have either just
// replayed the whole table up to h->step (lazy_rollover) or replaced the optimizer state wholesale (restore / import).
int lazy_rebase(ScoreModel* h) {
    if (h->cfg.adam_mode != SCORE_ADAM_LAZY) return SCORE_OK;
    CK(cudaStreamSynchronize(h->st));
    launch_fill_i32(h->st, h->last_step, h->dm.V, h->step);
    CK(cudaMemsetAsync(h->alpha_buf, 0, sizeof(float) * h->alpha_cap, h->st));
    h->hist_base = h->step;
    h->alpha_hist = h->alpha_buf - h->hist_base;
    drop_graphs(h);   // the captured launches carry the old window pointer
    CK(cudaStreamSynchronize(h->st));
    return SCORE_OK;
}
// the window is nearly full: replay the skipped steps of every row, then restart the window (tf.train.AdamOptimizer has
// no step limit; at 0.35 ms/step a window of 2^20 steps lasts ~6 minutes)
int lazy_rollover_if_needed(ScoreModel* h) {
    if (h->cfg.adam_mode != SCORE_ADAM_LAZY || (int64_t)h->step + 2 - h->hist_base < h->alpha_cap) return SCORE_OK;
    int rc = flush_lazy(h);
    if (rc) return rc;
    return lazy_rebase(h);
}

// step counter from the Adam slot variables of a checkpoint (beta1_power = 0.9^(step+1), beta2_power = 0.999^(step+1)):
// beta1_power while it is a normal float (it goes denormal near 830 steps), beta2_power after that
void recover_step(ScoreModel* h) {
    double s = 0.0;
    const float b1 = h->beta1_power, b2 = h->beta2_power;
    if (b1 > 1e-30f && b1 < 1.0f) s = log((double)b1) / log(0.9) - 1.0;
    else if (b2 > 0.f && b2 < 1.0f) s = log((double)b2) / log(0.999) - 1.0;
    else return;   // no usable slot value: keep the counter
    if (!(s >= 0.0)) s = 0.0;
    if (s > 2.0e9) s = 2.0e9;
    h->step = (int32_t)llround(s);
}

int finish_sync(ScoreModel* h, float* loss_out) {
    CK(cudaMemcpyAsync(h->loss_host, h->loss_dev, 2 * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(h->err_host, h->err_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    CK(cudaGetLastError());
    if (h->probes_on && h->last_mode == MODE_TRAIN) {
        for (int p = 0; p < PR_COUNT; ++p) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, h->pr_beg[p], h->pr_end[p]) == cudaSuccess) { h->pr_ms[p] += ms; h->pr_n[p]++; }
            else cudaGetLastError();
        }
    }
    if (loss_out) *loss_out = *h->loss_host;
    if (*h->err_host) {
        const int32_t code = *h->err_host;
        cudaMemsetAsync(h->err_flag, 0, sizeof(int32_t), h->st);
        if (code == 2) return fail(h, SCORE_ERR_ARG, "data-parallel exchange: a rank's unique-row count exceeds the block capacity");
        return fail(h, SCORE_ERR_ID_RANGE, "an id in the batch is outside [0, feature_size)");
    }
    return SCORE_OK;
}

// Result of a training step without waiting for its backward pass and update: wait for the result packet only.  Every
// later call on the handle is ordered behind the step by the stream, so nothing can observe the pre-update state.
int finish_early(ScoreModel* h, float* loss_out) {
    volatile int32_t* pk = reinterpret_cast<volatile int32_t*>(h->early_host);
    const int32_t want = h->hyper_host->seq;
    bool got = false;
    for (int64_t spin = 0; !got; ++spin) {
        got = pk[3] == want;
        if (got || (spin & 0xfff) != 0xfff) continue;
        // every few thousand polls: has the whole step finished (or failed) without delivering the packet?
        const cudaError_t q = cudaStreamQuery(h->st);
        if (q == cudaErrorNotReady) continue;
        got = pk[3] == want;
        if (!got) return finish_sync(h, loss_out);
    }
    __sync_synchronize();
    int32_t err = pk[2];
    float lv[2];
    int32_t bits0 = pk[0], bits1 = pk[1];
    memcpy(&lv[0], &bits0, 4); memcpy(&lv[1], &bits1, 4);
    if (loss_out) *loss_out = lv[0];
    h->loss_host[0] = lv[0]; h->loss_host[1] = lv[1];
    if (err) {
        cudaMemsetAsync(h->err_flag, 0, sizeof(int32_t), h->st);
        return fail(h, SCORE_ERR_ID_RANGE, "an id in the batch is outside [0, feature_size)");
    }
    return SCORE_OK;
}

int run_step(ScoreModel* h, const ScoreBatch* b, int mode, float lr, float reg_lambda, float keep_prob,
             int global_batch, bool sync, float* loss_out) {
    int rc = check_batch(h, b);
    if (rc) return rc;
    CK(cudaSetDevice(h->device));
    const int B = b->batch_size;
    rc = ensure_workspace(h, B > h->cfg.max_batch ? B : h->cfg.max_batch);
    if (rc) return rc;
    set_batch_dims(h, B);
    if (mode == MODE_FWDBWD) { rc = ensure_seg(h, h->dm.N); if (rc) return rc; }
    const bool train = (mode == MODE_TRAIN);
    if (train) { rc = lazy_rollover_if_needed(h); if (rc) return rc; }
    rc = upload_batch(h, b);
    if (rc) return rc;
    rc = upload_hyper(h, B, lr, reg_lambda, (mode == MODE_EVAL) ? 1.0f : keep_prob, mode != MODE_EVAL, global_batch);
    if (rc) return rc;
    h->last_N = h->dm.N;
    h->last_mode = mode;

    bool launched = false;
    if (train && h->cfg.use_graph) {
        auto it = h->graphs_train.find(B);
        if (it != h->graphs_train.end()) {
            CK(cudaGraphLaunch(it->second, h->st));
            g_launch_count += h->graph_kernels[B];
            launched = true;
        } else if (h->warm_train[B] >= 1) {
            // second time this batch size is seen: capture (the first run set all func attributes)
            cudaGraph_t graph = nullptr;
            h->capturing = true;
            const int64_t before = g_launch_count;
            CK(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
            enqueue_step(h, mode);
            h->graph_kernels[B] = g_launch_count - before;
            cudaError_t ce = cudaStreamEndCapture(h->st, &graph);
            h->capturing = false;
            if (ce != cudaSuccess) { h->err = std::string("graph capture failed: ") + cudaGetErrorString(ce); return SCORE_ERR_CUDA; }
            cudaGraphExec_t exec = nullptr;
            CK(cudaGraphInstantiate(&exec, graph, 0));
            cudaGraphDestroy(graph);
            h->graphs_train[B] = exec;
            CK(cudaGraphLaunch(exec, h->st));
            launched = true;
        }
        h->warm_train[B]++;
    }
    if (!launched) enqueue_step(h, mode);
    if (train) {
        h->step += 1;
        h->beta1_power = h->beta1_power * 0.9f;
        h->beta2_power = h->beta2_power * 0.999f;
    }
    stage_release(h);
    h->pending_step = true;
    if (sync) {
        if (train && !h->probes_on && early_return()) return finish_early(h, loss_out);
        h->pending_step = false;
        return finish_sync(h, loss_out);
    }
    return SCORE_OK;
}

}  // namespace

// ============================================================================================ C ABI
extern "C" {

const char* score_last_error(ScoreHandle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int score_create(const ScoreConfig* cfg, int device, ScoreHandle* out) {
    if (!cfg || !out) { g_create_error = "cfg/out is NULL"; return SCORE_ERR_ARG; }
    *out = nullptr;
    auto bad = [&](const char* m) { g_create_error = m; return SCORE_ERR_ARG; };
    if (cfg->feature_size < 2 || cfg->feature_size >= ((int64_t)1 << 31)) return bad("feature_size must be in [2, 2^31)");
    const int d = cfg->eb_dim;
    if (d < 4 || d > 128 || (d & (d - 1))) return bad("eb_dim must be a power of two in [4, 128]");
    if (cfg->obj_per_time_slice < 1 || cfg->obj_per_time_slice > 32) return bad("obj_per_time_slice must be in [1, 32]");
    if (cfg->hidden_size < 1 || cfg->hidden_size > 128) return bad("hidden_size must be in [1, 128]");
    if (cfg->max_time_len < 1 || cfg->user_fnum < 1 || cfg->item_fnum < 1) return bad("max_time_len / fnum must be positive");
    if (cfg->model_type < 0 || cfg->model_type > SCORE_MODEL_RRN) return bad("unknown model_type");
    if (cfg->adam_mode < 0 || cfg->adam_mode > 2) return bad("unknown adam_mode");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
        return SCORE_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) return bad("device index out of range");
    ScoreModel* h = new ScoreModel();
    h->cfg = *cfg;
    h->device = device;
    auto die = [&](int rc) { g_create_error = h->err; score_destroy(h); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) { h->err = "cudaSetDevice failed"; return die(SCORE_ERR_CUDA); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { h->err = "cudaGetDeviceProperties failed"; return die(SCORE_ERR_CUDA); }
    if (prop.major != 10) {
        h->err = "this build contains sm_100a kernels only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
        return die(SCORE_ERR_CUDA);
    }
    if (const char* e = getenv("SCORE_L2_FETCH")) {   // profiling knob: L2 fetch granularity hint (32 / 64 / 128 bytes)
        const int gran = atoi(e);
        if (gran > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        fprintf(stderr, "score_b200: cudaLimitMaxL2FetchGranularity = %zu\n", got);
    }
    // the main stream carries the critical path; the sort / weight-gradient branches only fill idle SMs
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&h->st, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->st2, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->st_w, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_l2, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_w, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_tgt, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_q, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_prep, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_keys, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fc, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_att, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->st_h2d, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_h2d[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_h2d[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_stage_free[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_stage_free[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_qb, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        h->err = "stream/event creation failed";
        return die(SCORE_ERR_CUDA);
    }
    for (int p = 0; p < PR_COUNT; ++p) {
        cudaEventCreate(&h->pr_beg[p]); cudaEventCreate(&h->pr_end[p]);
        h->pr_ms[p] = 0; h->pr_n[p] = 0;
    }
    for (int i = 0; i < 16; ++i) cudaEventCreateWithFlags(&h->ev_pool[i], cudaEventDisableTiming);
    build_registry(h);
    int rc = alloc_params(h);
    if (rc) return die(rc);
    if (cfg->init_weights) { rc = init_weights(h); if (rc) return die(rc); }
    if (cudaStreamSynchronize(h->st) != cudaSuccess) { h->err = "initialisation failed"; return die(SCORE_ERR_CUDA); }
    h->launches0 = g_launch_count;
    *out = h;
    return SCORE_OK;
}

int score_destroy(ScoreHandle h) {
    if (!h) return SCORE_OK;
    cudaSetDevice(h->device);
    if (h->st) cudaStreamSynchronize(h->st);
    free_workspace(h);
    for (void* p : {(void*)h->emb_tab, (void*)h->last_step, (void*)h->P, (void*)h->G,
                    (void*)h->M1, (void*)h->V1, (void*)h->PG, (void*)h->flags, (void*)h->Dv, (void*)h->n_heads_dev, (void*)h->claim_counter, (void*)h->claim_ext, (void*)h->alpha_buf, (void*)h->l2sum,
                    (void*)h->loss_dev, (void*)h->err_flag, (void*)h->step_dev, (void*)h->seg_rows, (void*)h->seg_heads,
                    (void*)h->sh_owner, (void*)h->sh_counts, (void*)h->sh_send_rows, (void*)h->sh_sel, (void*)h->sh_mini,
                    (void*)h->sh_staged, (void*)h->sh_grad_send})
        if (p) cudaFree(p);
    if (h->ev_presort) cudaEventDestroy(h->ev_presort);
    if (h->ev_sh_counts) cudaEventDestroy(h->ev_sh_counts);
    if (h->sh_counts_host) cudaFreeHost(h->sh_counts_host);
    if (h->hyper_ring) cudaFreeHost(h->hyper_ring);
    for (int i = 0; i < ScoreModel::kHyperSlots; ++i) if (h->hyper_ev[i]) cudaEventDestroy(h->hyper_ev[i]);
    if (h->loss_host) cudaFreeHost(h->loss_host);
    if (h->err_host) cudaFreeHost(h->err_host);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_keys) cudaEventDestroy(h->ev_keys);
    if (h->ev_fc) cudaEventDestroy(h->ev_fc);
    if (h->ev_att) cudaEventDestroy(h->ev_att);
    if (h->ev_qb) cudaEventDestroy(h->ev_qb);
    if (h->ev_l2) cudaEventDestroy(h->ev_l2);
    if (h->ev_w) cudaEventDestroy(h->ev_w);
    if (h->ev_tgt) cudaEventDestroy(h->ev_tgt);
    if (h->ev_q) cudaEventDestroy(h->ev_q);
    if (h->ev_prep) cudaEventDestroy(h->ev_prep);
    for (int i = 0; i < 16; ++i) if (h->ev_pool[i]) cudaEventDestroy(h->ev_pool[i]);
    if (h->st_w) cudaStreamDestroy(h->st_w);
    if (h->st) {
        for (int p = 0; p < PR_COUNT; ++p) { cudaEventDestroy(h->pr_beg[p]); cudaEventDestroy(h->pr_end[p]); }
        cudaStreamDestroy(h->st);
    }
    if (h->st2) cudaStreamDestroy(h->st2);
    if (h->st_h2d) cudaStreamDestroy(h->st_h2d);
    for (int i = 0; i < 2; ++i) { if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]); if (h->ev_stage_free[i]) cudaEventDestroy(h->ev_stage_free[i]); }
    if (h->early_host) cudaFreeHost(h->early_host);
    if (h->st_cnt) cudaStreamDestroy(h->st_cnt);
    if (h->ev_counts) cudaEventDestroy(h->ev_counts);
    if (h->cnt_slot) cudaFree(h->cnt_slot);
    if (h->cnt_host) cudaFreeHost(h->cnt_host);
    if (h->dp_block) cudaFree(h->dp_block);
    if (h->dp_sample) cudaFree(h->dp_sample);
    if (h->loss_glob) cudaFree(h->loss_glob);
    if (h->loss_glob_host) cudaFreeHost(h->loss_glob_host);
    delete h;
    return SCORE_OK;
}

int score_train_step(ScoreHandle h, const ScoreBatch* batch, float lr, float reg_lambda, float keep_prob,
                     float* loss_out) {
    if (!h) return SCORE_ERR_ARG;
    return run_step(h, batch, MODE_TRAIN, lr, reg_lambda, keep_prob, 0, true, loss_out);
}

int score_train_step_async(ScoreHandle h, const ScoreBatch* batch, float lr, float reg_lambda, float keep_prob) {
    if (!h) return SCORE_ERR_ARG;
    return run_step(h, batch, MODE_TRAIN, lr, reg_lambda, keep_prob, 0, false, nullptr);
}

int score_wait(ScoreHandle h, float* loss_out) {
    if (!h) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    h->pending_step = false;
    return finish_sync(h, loss_out);
}

int score_eval(ScoreHandle h, const ScoreBatch* batch, float reg_lambda, float* preds_out, float* loss_out) {
    if (!h) return SCORE_ERR_ARG;
    int rc = run_step(h, batch, MODE_EVAL, 0.f, reg_lambda, 1.0f, 0, false, nullptr);
    if (rc) return rc;
    if (preds_out) CK(cudaMemcpyAsync(preds_out, h->y, sizeof(float) * batch->batch_size, cudaMemcpyDeviceToHost, h->st));
    return finish_sync(h, loss_out);
}

int score_forward_backward(ScoreHandle h, const ScoreBatch* batch, float reg_lambda, float keep_prob, float* loss_out) {
    if (!h) return SCORE_ERR_ARG;
    return run_step(h, batch, MODE_FWDBWD, 0.f, reg_lambda, keep_prob, 0, true, loss_out);
}

int score_get_buffer(ScoreHandle h, const char* name, void* data, size_t capacity_bytes, size_t* count, int* dtype) {
    if (!h || !name) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    std::string n(name);
    const void* src = nullptr; size_t cnt = 0; int dt = 0;
    const Dims& dm = h->dm;
    if (n.rfind("grad/", 0) == 0) {
        const Tensor* t = find_tensor(h, n.substr(5));
        if (!t || t->is_emb) return fail(h, SCORE_ERR_NAME, "unknown gradient: " + n);
        src = h->G + t->off; cnt = (size_t)(t->rows * t->cols);
    } else if (n == "emb_grad/heads") {
        src = h->seg_heads; cnt = (size_t)h->last_N; dt = 1;
    } else if (n == "emb_grad/seg_rows") {
        src = h->seg_rows; cnt = (size_t)h->last_N * dm.d;
    } else if (n == "sorted_keys") {
        src = h->sb.keys[h->sort_out]; cnt = (size_t)h->last_N; dt = 1;
    } else if (n == "sorted_pos") {
        src = h->sb.vals[h->sort_out]; cnt = (size_t)h->last_N; dt = 1;
    } else {
        auto it = h->bufs.find(n);
        if (it == h->bufs.end()) return fail(h, SCORE_ERR_NAME, "unknown buffer: " + n);
        src = it->second.ptr; dt = it->second.dtype;
        // report the live extent for the current batch, not the capacity
        cnt = it->second.count / (size_t)h->cap_B * (size_t)dm.B;
    }
    if (count) *count = cnt;
    if (dtype) *dtype = dt;
    if (!data) return SCORE_OK;
    if (!src) return fail(h, SCORE_ERR_ARG, "buffer not populated: " + n);
    if (capacity_bytes < cnt * 4) return fail(h, SCORE_ERR_ARG, "capacity too small for " + n);
    CK(cudaMemcpyAsync(data, src, cnt * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return SCORE_OK;
}

int score_tensor_count(ScoreHandle h) { return h ? (int)h->tensors.size() : 0; }

int score_tensor_info(ScoreHandle h, int index, char* name, size_t name_cap, int64_t* rows, int64_t* cols) {
    if (!h || index < 0 || index >= (int)h->tensors.size()) return SCORE_ERR_ARG;
    const Tensor& t = h->tensors[index];
    if (name && name_cap) { strncpy(name, t.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (rows) *rows = t.rows;
    if (cols) *cols = t.cols;
    return SCORE_OK;
}

}  // extern "C"

namespace {

// resolve "<var>", "<var>/Adam", "<var>/Adam_1" to a device pointer + element count
// rows [row0, row0 + nrows) of one field (var / m / v) of the embedding records <-> contiguous host memory [nrows, d]
int emb_copy(ScoreModel* h, float* field, int64_t row0, int64_t nrows, float* host, bool to_host) {
    const size_t w = sizeof(float) * h->dm.d, pitch = sizeof(float) * h->es;
    const int64_t step = (int64_t)1 << 20;
    for (int64_t r = 0; r < nrows; r += step) {
        const int64_t n = nrows - r < step ? nrows - r : step;
        float* dev = field + (row0 + r) * h->es;
        float* hp = host + r * h->dm.d;
        if (to_host) CK(cudaMemcpy2DAsync(hp, w, dev, pitch, w, (size_t)n, cudaMemcpyDeviceToHost, h->st));
        else CK(cudaMemcpy2DAsync(dev, pitch, hp, w, w, (size_t)n, cudaMemcpyHostToDevice, h->st));
    }
    CK(cudaStreamSynchronize(h->st));
    return SCORE_OK;
}

int resolve(ScoreModel* h, const std::string& name, float** ptr, size_t* count, bool* is_emb) {
    std::string base = name;
    int slot = 0;
    auto ends = [&](const std::string& s, const char* suf) {
        size_t l = strlen(suf);
        return s.size() > l && s.compare(s.size() - l, l, suf) == 0;
    };
    if (ends(name, "/Adam_1")) { base = name.substr(0, name.size() - 7); slot = 2; }
    else if (ends(name, "/Adam")) { base = name.substr(0, name.size() - 5); slot = 1; }
    const Tensor* t = find_tensor(h, base);
    if (!t) return fail(h, SCORE_ERR_NAME, "unknown tensor: " + name);
    if (slot && !(t->flags & 2)) return fail(h, SCORE_ERR_NAME, "no Adam slot for non-trainable " + base);
    *count = (size_t)(t->rows * t->cols);
    *is_emb = t->is_emb;
    if (t->is_emb) *ptr = slot == 0 ? h->emb : slot == 1 ? h->emb_m : h->emb_v;
    else *ptr = (slot == 0 ? h->P : slot == 1 ? h->M1 : h->V1) + t->off;
    return SCORE_OK;
}

}  // namespace

extern "C" {

int score_get_tensor(ScoreHandle h, const char* name, float* data, size_t count) {
    if (!h || !name || !data) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    std::string n(name);
    if (n == "beta1_power" || n == "beta2_power") {
        if (count < 1) return fail(h, SCORE_ERR_ARG, "count too small");
        data[0] = n == "beta1_power" ? h->beta1_power : h->beta2_power;
        return SCORE_OK;
    }
    if (n == "step") {
        if (count < 1) return fail(h, SCORE_ERR_ARG, "count too small");
        data[0] = (float)h->step;
        return SCORE_OK;
    }
    float* p; size_t cnt; bool is_emb;
    int rc = resolve(h, n, &p, &cnt, &is_emb);
    if (rc) return rc;
    if (count != cnt) return fail(h, SCORE_ERR_ARG, "element count mismatch for " + n);
    if (is_emb) {
        rc = flush_lazy(h);
        if (rc) return rc;
        return emb_copy(h, p, 0, h->dm.V, data, true);
    }
    CK(cudaMemcpyAsync(data, p, cnt * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return SCORE_OK;
}

int score_set_tensor(ScoreHandle h, const char* name, const float* data, size_t count) {
    if (!h || !name || !data) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    std::string n(name);
    if (n == "beta1_power" || n == "beta2_power") {
        if (count < 1) return fail(h, SCORE_ERR_ARG, "count too small");
        (n == "beta1_power" ? h->beta1_power : h->beta2_power) = data[0];
        recover_step(h);
        return lazy_rebase(h);   // imported optimizer state is current at the imported step
    }
    if (n == "step") {   // explicit step counter (exact; the beta powers only approximate it after ~10^4 steps)
        if (count < 1 || !(data[0] >= 0.f)) return fail(h, SCORE_ERR_ARG, "step must be a non-negative number");
        h->step = (int32_t)llround((double)data[0]);
        return lazy_rebase(h);
    }
    float* p; size_t cnt; bool is_emb;
    int rc = resolve(h, n, &p, &cnt, &is_emb);
    if (rc) return rc;
    if (count != cnt) return fail(h, SCORE_ERR_ARG, "element count mismatch for " + n);
    if (is_emb) {
        rc = flush_lazy(h);
        if (rc) return rc;
        rc = emb_copy(h, p, 0, h->dm.V, const_cast<float*>(data), false);
        if (rc) return rc;
        if (h->last_step) {   // externally written state: the rows count as touched (current at the present step)
            launch_fill_i32(h->st, h->last_step, h->dm.V, h->step);
            CK(cudaStreamSynchronize(h->st));
        }
        return SCORE_OK;
    }
    CK(cudaMemcpyAsync(p, data, cnt * sizeof(float), cudaMemcpyHostToDevice, h->st));
    CK(cudaStreamSynchronize(h->st));
    return SCORE_OK;
}

int score_get_rows(ScoreHandle h, const char* name, int64_t row0, int64_t nrows, float* data) {
    if (!h || !name || !data) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    float* p; size_t cnt; bool is_emb;
    int rc = resolve(h, name, &p, &cnt, &is_emb);
    if (rc) return rc;
    if (!is_emb) return fail(h, SCORE_ERR_ARG, "row access is for emb_mtx and its slots");
    if (row0 < 0 || nrows < 0 || row0 + nrows > h->dm.V) return fail(h, SCORE_ERR_ARG, "row range out of bounds");
    rc = flush_lazy(h);
    if (rc) return rc;
    return emb_copy(h, p, row0, nrows, data, true);
}

int score_set_rows(ScoreHandle h, const char* name, int64_t row0, int64_t nrows, const float* data) {
    if (!h || !name || !data) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    float* p; size_t cnt; bool is_emb;
    int rc = resolve(h, name, &p, &cnt, &is_emb);
    if (rc) return rc;
    if (!is_emb) return fail(h, SCORE_ERR_ARG, "row access is for emb_mtx and its slots");
    if (row0 < 0 || nrows < 0 || row0 + nrows > h->dm.V) return fail(h, SCORE_ERR_ARG, "row range out of bounds");
    rc = flush_lazy(h);
    if (rc) return rc;
    rc = emb_copy(h, p, row0, nrows, const_cast<float*>(data), false);
    if (rc) return rc;
    if (h->last_step && nrows > 0) {   // externally written rows count as touched (see alloc_params: last_step)
        launch_fill_i32(h->st, h->last_step + row0, nrows, h->step);
        CK(cudaStreamSynchronize(h->st));
    }
    return SCORE_OK;
}

// Checkpoint: "SCB2CKPT" | version | n tensors | per tensor {name, rows, cols, var, [Adam, Adam_1]} | beta powers.
// Keys are the TF variable names so a tf.train.Saver user can convert (score.py:135-142 saves all global variables).
int score_save(ScoreHandle h, const char* path) {
    if (!h || !path) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    int rc = flush_lazy(h);
    if (rc) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(h, SCORE_ERR_IO, std::string("cannot open for writing: ") + path);
    const char magic[8] = {'S', 'C', 'B', '2', 'C', 'K', 'P', 'T'};
    uint32_t ver = 1, nt = (uint32_t)h->tensors.size();
    bool ok = fwrite(magic, 1, 8, f) == 8 && fwrite(&ver, 4, 1, f) == 1 && fwrite(&nt, 4, 1, f) == 1;
    const size_t chunk = (size_t)1 << 24;   // floats per staging chunk
    std::vector<float> host(chunk);
    for (const Tensor& t : h->tensors) {
        uint32_t nl = (uint32_t)t.name.size();
        uint32_t nslots = (t.flags & 2) ? 3 : 1;
        ok = ok && fwrite(&nl, 4, 1, f) == 1 && fwrite(t.name.data(), 1, nl, f) == nl && fwrite(&t.rows, 8, 1, f) == 1 &&
             fwrite(&t.cols, 8, 1, f) == 1 && fwrite(&nslots, 4, 1, f) == 1;
        const size_t cnt = (size_t)(t.rows * t.cols);
        for (uint32_t s = 0; s < nslots && ok; ++s) {
            const float* src = t.is_emb ? (s == 0 ? h->emb : s == 1 ? h->emb_m : h->emb_v)
                                        : (s == 0 ? h->P : s == 1 ? h->M1 : h->V1) + t.off;
            for (size_t o = 0; o < cnt && ok; o += chunk) {
                size_t n = cnt - o < chunk ? cnt - o : chunk;   // chunk is a multiple of every supported row width
                const bool read_ok = t.is_emb
                    ? emb_copy(h, const_cast<float*>(src), (int64_t)(o / h->dm.d), (int64_t)(n / h->dm.d), host.data(), true) == SCORE_OK
                    : cudaMemcpy(host.data(), src + o, n * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess;
                if (!read_ok) {
                    fclose(f);
                    return fail(h, SCORE_ERR_CUDA, "device read failed during save");
                }
                ok = fwrite(host.data(), sizeof(float), n, f) == n;
            }
        }
    }
    ok = ok && fwrite(&h->beta1_power, 4, 1, f) == 1 && fwrite(&h->beta2_power, 4, 1, f) == 1 &&
         fwrite(&h->step, 4, 1, f) == 1;
    ok = (fclose(f) == 0) && ok;
    return ok ? SCORE_OK : fail(h, SCORE_ERR_IO, std::string("short write: ") + path);
}

int score_restore(ScoreHandle h, const char* path) {
    if (!h || !path) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    FILE* f = fopen(path, "rb");
    if (!f) return fail(h, SCORE_ERR_IO, std::string("cannot open checkpoint: ") + path);
    char magic[8]; uint32_t ver = 0, nt = 0;
    bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "SCB2CKPT", 8) == 0 && fread(&ver, 4, 1, f) == 1 &&
              fread(&nt, 4, 1, f) == 1 && ver == 1;
    if (!ok || nt != h->tensors.size()) { fclose(f); return fail(h, SCORE_ERR_IO, "not a matching score_b200 checkpoint"); }
    const size_t chunk = (size_t)1 << 24;
    std::vector<float> host(chunk);
    for (uint32_t i = 0; i < nt; ++i) {
        uint32_t nl = 0, nslots = 0; int64_t rows = 0, cols = 0;
        if (fread(&nl, 4, 1, f) != 1 || nl > 4096) { fclose(f); return fail(h, SCORE_ERR_IO, "corrupt checkpoint"); }
        std::string name(nl, '\0');
        if (fread(&name[0], 1, nl, f) != nl || fread(&rows, 8, 1, f) != 1 || fread(&cols, 8, 1, f) != 1 ||
            fread(&nslots, 4, 1, f) != 1) { fclose(f); return fail(h, SCORE_ERR_IO, "corrupt checkpoint"); }
        const Tensor* t = find_tensor(h, name);
        if (!t || t->rows != rows || t->cols != cols) { fclose(f); return fail(h, SCORE_ERR_IO, "checkpoint tensor mismatch: " + name); }
        const size_t cnt = (size_t)(rows * cols);
        for (uint32_t s = 0; s < nslots; ++s) {
            float* dst = t->is_emb ? (s == 0 ? h->emb : s == 1 ? h->emb_m : h->emb_v)
                                   : (s == 0 ? h->P : s == 1 ? h->M1 : h->V1) + t->off;
            for (size_t o = 0; o < cnt; o += chunk) {
                size_t n = cnt - o < chunk ? cnt - o : chunk;
                if (fread(host.data(), sizeof(float), n, f) != n) { fclose(f); return fail(h, SCORE_ERR_IO, "truncated checkpoint"); }
                const bool write_ok = t->is_emb
                    ? emb_copy(h, dst, (int64_t)(o / h->dm.d), (int64_t)(n / h->dm.d), host.data(), false) == SCORE_OK
                    : cudaMemcpy(dst + o, host.data(), n * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
                if (!write_ok) {
                    fclose(f);
                    return fail(h, SCORE_ERR_CUDA, "device write failed during restore");
                }
            }
        }
    }
    ok = fread(&h->beta1_power, 4, 1, f) == 1 && fread(&h->beta2_power, 4, 1, f) == 1 && fread(&h->step, 4, 1, f) == 1;
    fclose(f);
    if (!ok) return fail(h, SCORE_ERR_IO, "truncated checkpoint");
    if (h->cfg.adam_mode == SCORE_ADAM_LAZY) return lazy_rebase(h);   // every row of a restored table is current at the restored step
    if (h->last_step) {
        launch_fill_i32(h->st, h->last_step, h->dm.V, h->step);
        CK(cudaStreamSynchronize(h->st));
    }
    return SCORE_OK;
}

int score_eval_metrics(ScoreHandle h, const float* preds, const int32_t* target_iids, const int32_t* labels,
                       int64_t n, int32_t group, double* out9) {
    if (!h || !preds || !target_iids || !labels || !out9) return SCORE_ERR_ARG;
    if (n <= 0 || group <= 0 || n % group != 0) return fail(h, SCORE_ERR_ARG, "n must be a positive multiple of group");
    if (n >= ((int64_t)1 << 31)) return fail(h, SCORE_ERR_ARG, "n too large");
    CK(cudaSetDevice(h->device));
    float* d_preds = nullptr; int32_t *d_iids = nullptr, *d_labels = nullptr, *d_keys = nullptr, *d_rank = nullptr;
    double *d_terms = nullptr, *d_sums = nullptr;
    SortBufs sb{};
    uint32_t* hist = nullptr;
    std::vector<void*> owned;
    auto cleanup = [&]() { for (void* p : owned) cudaFree(p); };
    auto A = [&](void** p, size_t bytes) { cudaError_t e = cudaMalloc(p, bytes); if (e == cudaSuccess) owned.push_back(*p); return e; };
    const int64_t n_groups = n / group;
    const size_t nterms = (size_t)(2 * n > 6 * n_groups ? 2 * n : 6 * n_groups);
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = A((void**)&d_preds, sizeof(float) * n);
    if (e == cudaSuccess) e = A((void**)&d_iids, sizeof(int32_t) * n);
    if (e == cudaSuccess) e = A((void**)&d_labels, sizeof(int32_t) * n);
    if (e == cudaSuccess) e = A((void**)&d_keys, sizeof(int32_t) * n);
    if (e == cudaSuccess) e = A((void**)&d_rank, sizeof(int32_t) * n_groups);
    if (e == cudaSuccess) e = A((void**)&d_terms, sizeof(double) * nterms);
    if (e == cudaSuccess) e = A((void**)&d_sums, sizeof(double) * 16);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = A((void**)&sb.keys[i], sizeof(int32_t) * n);
        if (e == cudaSuccess) e = A((void**)&sb.vals[i], sizeof(int32_t) * n);
    }
    if (e == cudaSuccess) e = A((void**)&hist, sizeof(uint32_t) * sort_hist_elems(n));
    sb.hist = hist;
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_preds, preds, sizeof(float) * n, cudaMemcpyHostToDevice, h->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_iids, target_iids, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_labels, labels, sizeof(int32_t) * n, cudaMemcpyHostToDevice, h->st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_sums, 0, sizeof(double) * 16, h->st);
    if (e == cudaSuccess) e = compute_eval_metrics(h->st, d_preds, d_iids, d_labels, n, group, sb, d_keys, d_rank, d_terms, d_sums);
    double sums[16] = {0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(sums, d_sums, sizeof(sums), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    cleanup();
    if (e != cudaSuccess) return fail(h, SCORE_ERR_CUDA, std::string("eval metrics failed: ") + cudaGetErrorString(e));
    const double npos = sums[9], nneg = (double)n - npos;
    out9[0] = sums[0] / (double)n;
    out9[1] = (npos > 0 && nneg > 0) ? (sums[8] - npos * (npos + 1.0) * 0.5) / (npos * nneg) : NAN;
    for (int k = 0; k < 6; ++k) out9[2 + k] = sums[2 + k] / (double)n_groups;
    out9[8] = 0.0;
    return SCORE_OK;
}


}  // extern "C"

// ------------------------------------------------------------------------------------------ multi-GPU split step
namespace {

int ensure_ext_sort(ScoreModel* h, int64_t n) {
    if (n <= h->sb_ext_cap) return SCORE_OK;
    for (int i = 0; i < 2; ++i) { if (h->sb_ext.keys[i]) cudaFree(h->sb_ext.keys[i]); if (h->sb_ext.vals[i]) cudaFree(h->sb_ext.vals[i]); }
    if (h->sb_ext.hist) cudaFree(h->sb_ext.hist);
    if (h->sb_ext.runs) cudaFree(h->sb_ext.runs);
    if (h->sb_ext.runs_long) cudaFree(h->sb_ext.runs_long);
    if (h->sb_ext.part) cudaFree(h->sb_ext.part);
    if (h->sb_ext.slotinfo) cudaFree(h->sb_ext.slotinfo);
    if (h->sb_ext.done) cudaFree(h->sb_ext.done);
    int64_t cap = n + n / 4 + 1024;
    for (int i = 0; i < 2; ++i) {
        CK(cudaMalloc(&h->sb_ext.keys[i], sizeof(int32_t) * cap));
        CK(cudaMalloc(&h->sb_ext.vals[i], sizeof(int32_t) * cap));
    }
    CK(cudaMalloc(&h->sb_ext.hist, sizeof(uint32_t) * sort_hist_elems(cap)));
    CK(cudaMalloc(&h->sb_ext.runs, sizeof(int32_t) * 8 * cap));
    CK(cudaMalloc(&h->sb_ext.runs_long, sizeof(int32_t) * 4 * emb_runs_long_cap(cap)));
    CK(cudaMalloc(&h->sb_ext.part, sizeof(float) * emb_runs_part_cap(cap) * h->dm.d));
    CK(cudaMalloc(&h->sb_ext.slotinfo, sizeof(int32_t) * 4 * emb_runs_part_cap(cap)));
    CK(cudaMalloc(&h->sb_ext.done, sizeof(int32_t) * emb_runs_part_cap(cap)));
    CK(cudaMemsetAsync(h->sb_ext.done, 0, sizeof(int32_t) * emb_runs_part_cap(cap), h->st));
    h->sb_ext_cap = cap;
    return SCORE_OK;
}

// streams / scalars of the packed data-parallel exchange (created on first use)
int ensure_dp(ScoreModel* h) {
    if (h->st_cnt) return SCORE_OK;
    CK(cudaStreamCreateWithFlags(&h->st_cnt, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&h->ev_counts, cudaEventDisableTiming));
    CK(cudaMalloc(&h->cnt_slot, 2 * sizeof(int32_t)));
    CK(cudaMemset(h->cnt_slot, 0xff, 2 * sizeof(int32_t)));
    CK(cudaMallocHost(&h->cnt_host, 2 * sizeof(int32_t)));
    CK(cudaMalloc(&h->loss_glob, sizeof(double)));
    CK(cudaMemset(h->loss_glob, 0, sizeof(double)));
    CK(cudaMallocHost(&h->loss_glob_host, sizeof(double)));
    return SCORE_OK;
}

DpLayout dp_layout(const ScoreModel* h, const void* base, int world, int64_t cap) {
    DpLayout L{};
    L.base = static_cast<const int32_t*>(base); L.world = world; L.d = h->dm.d; L.cap = cap;
    L.dense_off = 128;
    L.keys_off = L.dense_off + (h->n_dense + 127) / 128 * 128;
    L.stride = L.keys_off + cap + cap * h->dm.d;
    return L;
}

}  // namespace

extern "C" {

// Upload a batch and build its sanitized key list (ids of masked slices -> 0, range check) without running the
// model; afterwards score_device_buffer(h, "keys") is valid.  First stage of a row-sharded step.
int score_prepare_batch(ScoreHandle h, const ScoreBatch* batch) {
    if (!h) return SCORE_ERR_ARG;
    int rc = check_batch(h, batch);
    if (rc) return rc;
    CK(cudaSetDevice(h->device));
    const int B = batch->batch_size;
    rc = ensure_workspace(h, B > h->cfg.max_batch ? B : h->cfg.max_batch);
    if (rc) return rc;
    set_batch_dims(h, B);
    rc = upload_batch(h, batch);
    if (rc) return rc;
    h->last_N = h->dm.N;
    Dims dm = h->dm;
    dm.V = ((int64_t)1 << 31) - 1;   // ids are GLOBAL row numbers here; the owner checks the range
    // carries the batch's pointer table; a begun step that still waits for score_step_finish keeps its hyper-parameters
    rc = h->begun ? upload_bp_keep_hyper(h) : upload_hyper(h, B, 0.f, 0.f, 1.f, 0, 0);
    if (rc) return rc;
    launch_build_keys(h->st, dm, h->bp_dev, h->keys, h->label, h->length, h->err_flag, nullptr, h->live);
    stage_release(h);
    return SCORE_OK;
}

// Device pointers of per-step buffers for the host-side exchange (torch.distributed on score_stream()):
// "keys" int32 [N], "grad_rows" float [N,d], "dense_grad" float [n_dense], "y_pred" float [B].
int score_device_buffer(ScoreHandle h, const char* name, void** dev_ptr, size_t* count) {
    if (!h || !name || !dev_ptr || !count) return SCORE_ERR_ARG;
    std::string n(name);
    if (n == "keys") { *dev_ptr = h->keys; *count = (size_t)h->last_N; }
    else if (n == "grad_rows") { *dev_ptr = h->grad_rows; *count = (size_t)h->last_N * h->dm.d; }
    else if (n == "dense_grad") { *dev_ptr = h->G; *count = (size_t)h->n_dense; }
    else if (n == "y_pred") { *dev_ptr = h->y; *count = (size_t)h->dm.B; }
    else return fail(h, SCORE_ERR_NAME, "unknown device buffer: " + n);
    return SCORE_OK;
}

// Owner side of a row-sharded gather: out[i] = current value of local row idx[i] (idx 0 -> zeros).  In LAZY mode the
// requested rows are first brought up to date.  All pointers are device pointers; ordered on score_stream().
namespace {
// LAZY: bring the rows idx_dev[0..n) (owner-local row numbers) up to date before they are served
int serve_catchup(ScoreModel* h, const int32_t* idx_dev, int64_t n) {
    int rc = upload_hyper(h, 1, 0.f, 0.f, 1.f, 0, 1);
    if (rc) return rc;
    if (h->cfg.adam_mode != SCORE_ADAM_LAZY) return SCORE_OK;
    if (n > h->claim_ext_cap) {
        CK(cudaStreamSynchronize(h->st));
        if (h->claim_ext) cudaFree(h->claim_ext);
        h->claim_ext = nullptr; h->claim_ext_cap = 0;
        CK(cudaMalloc(&h->claim_ext, sizeof(int32_t) * 2 * (n + n / 4)));
        h->claim_ext_cap = n + n / 4;
    }
    launch_emb_catchup_rows(h->st, idx_dev, n, h->dm.V, h->emb, h->emb_m, h->emb_v, h->last_step, h->dm.d, h->es, h->alpha_hist,
                            h->hyper_dev, h->claim_ext, h->claim_counter);
    return SCORE_OK;
}
}  // namespace

int score_gather_rows(ScoreHandle h, const int32_t* idx_dev, int64_t n, float* out_dev) {
    if (!h || (n > 0 && (!idx_dev || !out_dev))) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (n == 0) return SCORE_OK;
    int rc = serve_catchup(h, idx_dev, n);
    if (rc) return rc;
    launch_gather_rows(h->st, h->emb, h->es, idx_dev, n, h->dm.d, h->dm.V, out_dev, h->err_flag);
    return SCORE_OK;
}

// Peer-memory variants (NVLink): the owner stores the rows it serves straight into the requesters' staged tables, the
// requester stores its gradient rows straight into the owners' gradient buffers - gather / pack fused with the
// all-to-all.  count_matrix_dev: the all-gathered [world][world+1] counts; peers[r]: rank r's staged table (gradient
// buffer) as mapped into this process.  The caller runs a barrier across the ranks on score_stream() afterwards.
int score_shard_serve_push(ScoreHandle h, const int32_t* want_dev, int64_t n_recv, const int32_t* count_matrix_dev,
                           int32_t world, int32_t rank, const uint64_t* peers) {
    if (!h || !count_matrix_dev || !peers || world < 1 || world > SHARD_MAX_PEERS || rank < 0 || rank >= world || n_recv < 0 ||
        (n_recv > 0 && !want_dev))
        return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (n_recv > 0) {
        int rc = serve_catchup(h, want_dev, n_recv);
        if (rc) return rc;
    }
    ShardPeers sp{};
    for (int r = 0; r < world; ++r) sp.p[r] = reinterpret_cast<float*>(peers[r]);
    launch_shard_serve_push(h->st, h->emb, h->es, h->dm.d, h->dm.V, want_dev, n_recv, count_matrix_dev, world, rank, sp, h->err_flag);
    CK(cudaGetLastError());
    return SCORE_OK;
}
int score_shard_grad_push(ScoreHandle h, const int32_t* count_matrix_dev, int32_t world, int32_t rank, const uint64_t* peers) {
    if (!h || !count_matrix_dev || !peers || world < 1 || world > SHARD_MAX_PEERS || rank < 0 || rank >= world || !h->sh_world)
        return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    ShardPeers sp{};
    for (int r = 0; r < world; ++r) sp.p[r] = reinterpret_cast<float*>(peers[r]);
    launch_shard_grad_push(h->st, h->grad_rows, h->sh_sel, h->last_N, h->dm.d, count_matrix_dev, world, rank, sp);
    CK(cudaGetLastError());
    return SCORE_OK;
}
// staged tables that live outside the handle (symmetric memory of the peer-memory exchange): their addresses are stable,
// so the half-step on them may be captured as a CUDA graph like the handle's own staged table
int score_shard_register_staged(ScoreHandle h, const float* a, const float* b) {
    if (!h) return SCORE_ERR_ARG;
    if (a != h->sh_ext_staged[0] || b != h->sh_ext_staged[1]) {
        // the caller re-allocated its exchange buffers (they grow with the count matrix, which may happen while this
        // rank's own position count - and with it sh_cap - stays put): a half-step captured on the old tables must go
        CK(cudaSetDevice(h->device));
        CK(cudaStreamSynchronize(h->st));
        drop_graphs(h);
    }
    h->sh_ext_staged[0] = a; h->sh_ext_staged[1] = b;
    return SCORE_OK;
}

int score_shard_plan(ScoreHandle h, int32_t world, ScoreShardPlan* out) {
    if (!h || !out || world < 1 || world > 64) return SCORE_ERR_ARG;
    if (h->dm.B <= 0 || h->last_N <= 0) return fail(h, SCORE_ERR_ARG, "score_shard_plan needs score_prepare_batch first");
    CK(cudaSetDevice(h->device));
    const int64_t N = h->last_N;
    const int d = h->dm.d;
    if (N > h->sh_cap) {
        CK(cudaStreamSynchronize(h->st));
        drop_graphs(h);   // a captured half-step holds the old staged / mini_keys addresses
        cudaFree(h->sh_owner); cudaFree(h->sh_send_rows); cudaFree(h->sh_sel); cudaFree(h->sh_mini);
        cudaFree(h->sh_staged); cudaFree(h->sh_grad_send);
        const int64_t cap = N + N / 8;
        CK(cudaMalloc(&h->sh_owner, sizeof(int32_t) * cap));
        CK(cudaMalloc(&h->sh_send_rows, sizeof(int32_t) * cap));
        CK(cudaMalloc(&h->sh_sel, sizeof(int32_t) * cap));
        CK(cudaMalloc(&h->sh_mini, sizeof(int32_t) * cap));
        CK(cudaMalloc(&h->sh_staged, sizeof(float) * (cap + 1) * d));
        CK(cudaMemsetAsync(h->sh_staged, 0, sizeof(float) * d, h->st));
        CK(cudaMalloc(&h->sh_grad_send, sizeof(float) * cap * d));
        h->sh_cap = cap;
    }
    if (!h->sh_counts) CK(cudaMalloc(&h->sh_counts, sizeof(int32_t) * 65));
    h->sh_world = world;
    launch_shard_plan(h->st, h->sb, h->keys, N, world, h->sh_owner, h->sh_counts, h->sh_send_rows, h->sh_sel, h->sh_mini);
    out->counts = h->sh_counts; out->send_rows = h->sh_send_rows; out->staged = h->sh_staged; out->mini_keys = h->sh_mini;
    out->grad_send = h->sh_grad_send; out->n_positions = N;
    CK(cudaGetLastError());
    return SCORE_OK;
}

// Read-back of the all-gathered count matrix without draining the stream: the copy and its event are enqueued, the caller
// enqueues more work (the previous step's score_step_finish) and then waits for the event alone.
int score_shard_counts_fetch(ScoreHandle h, const int32_t* counts_dev, int32_t n) {
    if (!h || !counts_dev || n < 1 || n > 65 * 64) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (!h->sh_counts_host) {
        CK(cudaMallocHost(&h->sh_counts_host, sizeof(int32_t) * 65 * 64));
        CK(cudaEventCreateWithFlags(&h->ev_sh_counts, cudaEventDisableTiming));
    }
    CK(cudaMemcpyAsync(h->sh_counts_host, counts_dev, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->st));
    CK(cudaEventRecord(h->ev_sh_counts, h->st));
    return SCORE_OK;
}
int score_shard_counts_wait(ScoreHandle h, int32_t* out, int32_t n) {
    if (!h || !out || !h->sh_counts_host || n < 1 || n > 65 * 64) return SCORE_ERR_ARG;
    CK(cudaEventSynchronize(h->ev_sh_counts));
    memcpy(out, h->sh_counts_host, sizeof(int32_t) * n);
    return SCORE_OK;
}

int score_shard_pack_grads(ScoreHandle h) {
    if (!h || !h->sh_world) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    launch_shard_pack_grads(h->st, h->grad_rows, h->sh_sel, h->sh_counts, h->sh_world, h->last_N, h->dm.d, h->sh_grad_send);
    CK(cudaGetLastError());
    return SCORE_OK;
}

// Owner-side key list of a row-sharded step: known as soon as the ids have been exchanged, so its sort (and the run
// descriptors) run on the side stream under forward / backward instead of in front of the update.
int score_shard_presort(ScoreHandle h, const int32_t* ext_keys, int64_t n_ext) {
    if (!h || n_ext < 0 || (n_ext > 0 && !ext_keys)) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    h->presorted_keys = nullptr; h->presorted_n = 0;
    if (n_ext == 0) return SCORE_OK;
    if (n_ext >= ((int64_t)1 << 31)) return fail(h, SCORE_ERR_ARG, "too many gradient rows for int32 positions");
    if (n_ext > h->sb_ext_cap) CK(cudaStreamSynchronize(h->st));   // the previous step's update may still read the old buffers
    int rc = ensure_ext_sort(h, n_ext);
    if (rc) return rc;
    if (!h->ev_presort) CK(cudaEventCreateWithFlags(&h->ev_presort, cudaEventDisableTiming));
    cudaEvent_t e = h->ev_pool[h->ev_next++ & 15];
    CK(cudaEventRecord(e, h->st));                 // the key list is complete (the exchange ran on the main stream)
    CK(cudaStreamWaitEvent(h->st2, e, 0));
    const int out = launch_sort_pairs(h->st2, h->sb_ext, ext_keys, n_ext, key_bits(h->dm.V));
    launch_emb_runs(h->st2, h->sb_ext.keys[out], h->sb_ext.vals[out], n_ext, h->sb_ext.runs, h->sb_ext.runs_long, h->n_heads_dev, h->sb_ext.slotinfo);
    CK(cudaEventRecord(h->ev_presort, h->st2));
    h->presorted_keys = ext_keys; h->presorted_n = n_ext; h->presorted_out = out;
    CK(cudaGetLastError());
    return SCORE_OK;
}

// Forward + backward of one local batch WITHOUT the optimizer update; the loss mean uses 1/global_batch.
// batch == NULL reuses the batch of score_prepare_batch.  staged_table / staged_keys (device, may be NULL)
// replace the handle's table and key list for the gathers: staged_keys[p] indexes staged_table rows.
// Afterwards "dense_grad" and "grad_rows" hold this rank's contributions.
int score_step_begin(ScoreHandle h, const ScoreBatch* batch, float lr, float reg_lambda, float keep_prob,
                     int32_t global_batch, int32_t train, const float* staged_table, const int32_t* staged_keys) {
    if (!h) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (batch) {
        int rc = check_batch(h, batch);
        if (rc) return rc;
        const int B = batch->batch_size;
        rc = ensure_workspace(h, B > h->cfg.max_batch ? B : h->cfg.max_batch);
        if (rc) return rc;
        set_batch_dims(h, B);
        rc = upload_batch(h, batch);
        if (rc) return rc;
        h->last_N = h->dm.N;
    } else if (h->dm.B <= 0) {
        return fail(h, SCORE_ERR_ARG, "no prepared batch");
    }
    if (train) { int rc = lazy_rollover_if_needed(h); if (rc) return rc; }
    const Dims& dm = h->dm;
    {
        int rc = upload_hyper(h, dm.B, lr, reg_lambda, train ? keep_prob : 1.0f, train, global_batch);
        if (rc) return rc;
    }
    h->last_mode = MODE_BEGIN;
    h->begin_seq = h->hyper_host->seq;
    if (train && !staged_table) { int rc = ensure_dp(h); if (rc) return rc; }
    const bool with_keys = batch != nullptr;
    auto enqueue_begin = [&]() {
        const bool own_sort = train && !staged_table;   // sort of this rank's own keys under forward/backward (score_dp_pack)
        h->sort_dp = own_sort;
        h->early_on = false;
        h->sort_deferred = own_sort && sort_after_gather();
        cudaEventRecord(h->ev_fork, h->st);
        const bool lazy = h->cfg.adam_mode == SCORE_ADAM_LAZY;
        const bool fused_claim = with_keys && !staged_table && lazy;   // claim of the stale rows rides on build_keys
        if (with_keys) {
            ClaimArgs ca{h->last_step, h->hyper_dev, h->claim_list, h->claim_counter, 1};
            launch_build_keys(h->st, dm, h->bp_dev, h->keys, h->label, h->length, h->err_flag, fused_claim ? &ca : nullptr, h->live);
        }
        if (own_sort) {
            cudaEventRecord(h->ev_keys, h->st);
            enqueue_sort_branch(h, h->ev_keys, h->sort_deferred ? 1 : 0);
        }
        if (staged_table) {
            h->emb_fwd = staged_table; h->es_fwd = h->dm.d; h->keys_fwd = staged_keys;
        } else {
            h->emb_fwd = h->emb; h->es_fwd = h->es; h->keys_fwd = h->keys;
            if (fused_claim)
                launch_emb_replay(h->st, h->claim_list, h->claim_counter, dm.N, h->emb, h->emb_m, h->emb_v, dm.d, h->es, h->alpha_hist,
                                  h->hyper_dev, 1);
            else if (lazy)
                launch_emb_catchup_rows(h->st, h->keys, dm.N, dm.V, h->emb, h->emb_m, h->emb_v, h->last_step, dm.d, h->es,
                                        h->alpha_hist, h->hyper_dev, h->claim_list, h->claim_counter);
        }
        enqueue_forward(h, train != 0);
        if (train) enqueue_backward(h);
        if (train && !staged_table) cudaStreamWaitEvent(h->st, h->ev_join, 0);   // joins the sort branch (capture needs it)
    };
    h->local_sorted = train && !staged_table;
    // the training half-step is the same launch sequence every step: replay it as a CUDA graph (data-parallel: keyed by
    // B; row-sharded: only with the handle's own staged table / mini keys, whose addresses are stable - score_shard_plan)
    bool launched = false;
    const int staged_which = !staged_table ? -1 : staged_table == h->sh_staged ? 0 : staged_table == h->sh_ext_staged[0] ? 1
                             : staged_table == h->sh_ext_staged[1] ? 2 : -1;
    const bool staged_own = staged_which >= 0 && !with_keys && staged_keys == h->sh_mini;
    if (h->cfg.use_graph && train && ((!staged_table && with_keys) || staged_own)) {
        const int B = dm.B + (staged_own ? ((staged_which + 1) << 28) : 0);
        auto it = h->graphs_begin.find(B);
        if (it != h->graphs_begin.end()) {
            CK(cudaGraphLaunch(it->second, h->st));
            g_launch_count += h->graph_kernels_begin[B];
            if (staged_own) { h->emb_fwd = staged_table; h->es_fwd = h->dm.d; h->keys_fwd = staged_keys; }
            else { h->emb_fwd = h->emb; h->es_fwd = h->es; h->keys_fwd = h->keys; }
            launched = true;
        } else if (h->warm_begin[B] >= 1) {
            cudaGraph_t graph = nullptr;
            h->capturing = true;
            const int64_t before = g_launch_count;
            CK(cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal));
            enqueue_begin();
            h->graph_kernels_begin[B] = g_launch_count - before;
            cudaError_t ce = cudaStreamEndCapture(h->st, &graph);
            h->capturing = false;
            if (ce != cudaSuccess) { h->err = std::string("graph capture failed: ") + cudaGetErrorString(ce); return SCORE_ERR_CUDA; }
            cudaGraphExec_t exec = nullptr;
            CK(cudaGraphInstantiate(&exec, graph, 0));
            cudaGraphDestroy(graph);
            h->graphs_begin[B] = exec;
            CK(cudaGraphLaunch(exec, h->st));
            launched = true;
        }
        h->warm_begin[B]++;
    }
    if (!launched) enqueue_begin();
    stage_release(h);
    h->begun = train != 0;
    h->begun_lr = lr;
    CK(cudaGetLastError());
    return SCORE_OK;
}

// ---- data-parallel, replicated table: the packed exchange (kernels.h: DpLayout) --------------------------------
// Number of unique non-zero ids of the batch score_step_begin was given.  The count comes from the sort branch, which
// finishes long before forward/backward: the host waits for THAT (own stream + event), not for the main stream, so the
// device never idles while the exchange is being sized.
int score_dp_local_count(ScoreHandle h, int32_t* count_out) {
    if (!h || !count_out) return SCORE_ERR_ARG;
    if (!h->begun || !h->local_sorted) return fail(h, SCORE_ERR_ARG, "score_dp_local_count needs a training score_step_begin on the handle's own table");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamWaitEvent(h->st_cnt, h->ev_counts, 0));
    CK(cudaMemcpyAsync(h->cnt_host, h->cnt_slot, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->st_cnt));
    CK(cudaStreamSynchronize(h->st_cnt));
    if (h->cnt_host[1] != h->begin_seq) {
        // the event did not cover this step's launch (should not happen): fall back to draining the main stream
        if (!h->warned_stale) { fprintf(stderr, "score_b200: early count not ready (seq %d != %d), synchronising the step\n", h->cnt_host[1], h->begin_seq); h->warned_stale = true; }
        CK(cudaStreamSynchronize(h->st));
        CK(cudaMemcpy(h->cnt_host, h->cnt_slot, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (h->cnt_host[1] != h->begin_seq) return fail(h, SCORE_ERR_CUDA, "unique-row count of the step is missing");
    }
    *count_out = h->cnt_host[0];
    return SCORE_OK;
}

// words (4 bytes each) of one rank's block for a list capacity of `cap` rows
int64_t score_dp_block_words(ScoreHandle h, int64_t cap) {
    if (!h || cap <= 0) return 0;
    return dp_layout(h, nullptr, 1, cap).stride;
}

// Assemble this rank's block (device memory owned by the handle, valid until the next call): header, dense gradient,
// and its embedding gradient reduced to ONE row per unique id (deterministic segment reduce of scatter.cu in export
// mode), ascending by id.  cap: capacity of the id / row lists, a multiple of 1024 and >= every rank's count (all ranks
// must use the same value).  Call between score_step_begin and score_dp_finish; the all-gather of the blocks is the
// caller's (NCCL on score_stream()).
int score_dp_pack(ScoreHandle h, int64_t cap, void** block_dev, int64_t* block_words) {
    if (!h || !block_dev || !block_words) return SCORE_ERR_ARG;
    if (!h->begun || !h->local_sorted) return fail(h, SCORE_ERR_ARG, "score_dp_pack needs a training score_step_begin on the handle's own table");
    if (cap <= 0 || cap % 1024) return fail(h, SCORE_ERR_ARG, "cap must be a positive multiple of 1024");
    CK(cudaSetDevice(h->device));
    const Dims& dm = h->dm;
    const DpLayout L = dp_layout(h, nullptr, 1, cap);
    if (L.stride > h->dp_block_words) {
        CK(cudaStreamSynchronize(h->st));
        if (h->dp_block) cudaFree(h->dp_block);
        h->dp_block = nullptr; h->dp_block_words = 0;
        const int64_t words = L.stride + L.stride / 4;
        CK(cudaMalloc(&h->dp_block, sizeof(int32_t) * words));
        CK(cudaMemsetAsync(h->dp_block, 0, sizeof(int32_t) * words, h->st));
        h->dp_block_words = words;
    }
    int32_t* blk = h->dp_block;
    // the sort branch and the loss branch have been joined into the main stream by score_step_begin
    launch_dp_pack_misc(h->st, blk, L, h->n_heads_dev, h->hyper_dev, h->loss_dev, h->G, (int)h->n_dense);
    EmbUpdateArgs ea{};
    ea.skeys = h->sb.keys[h->sort_out]; ea.spos = h->sb.vals[h->sort_out]; ea.n = dm.N;
    ea.runs = h->sb.runs; ea.runs_long = h->sb.runs_long; ea.long_cap = emb_runs_long_cap(dm.N); ea.counters = h->n_heads_dev;
        ea.part = h->sb.part; ea.slotinfo = h->sb.slotinfo; ea.done = h->sb.done;
    ea.grad_rows = h->grad_rows; ea.d = dm.d; ea.hp = h->hyper_dev; ea.mode = 2;
    ea.out_heads = blk + L.keys_off; ea.out_rows = reinterpret_cast<float*>(blk + L.keys_off + cap);
    ea.head_slot = h->head_slot; ea.out_cap = cap;
    launch_emb_update(h->st, ea);
    *block_dev = blk; *block_words = L.stride;
    CK(cudaGetLastError());
    return SCORE_OK;
}

// Opt-in peer-memory exchange: store the block score_dp_pack assembled into every replica's gathered buffer.
// peer_bases[r] = device address (mapped into this process, e.g. by torch's symmetric memory) of replica r's buffer;
// the block lands at word offset dst_off_words in each of them.  The caller orders the stores against the readers
// (a barrier across the replicas on score_stream()) before score_dp_finish runs on its own buffer.
int score_dp_push(ScoreHandle h, int64_t cap, const uint64_t* peer_bases, int32_t world, int64_t dst_off_words) {
    if (!h || !peer_bases || world < 1 || world > DP_MAX_WORLD) return SCORE_ERR_ARG;
    if (!h->begun || !h->dp_block) return fail(h, SCORE_ERR_ARG, "score_dp_push needs score_dp_pack");
    CK(cudaSetDevice(h->device));
    const DpLayout L = dp_layout(h, nullptr, 1, cap);
    if (L.stride > h->dp_block_words || dst_off_words < 0 || dst_off_words % 4) return fail(h, SCORE_ERR_ARG, "bad block / offset");
    DpPeers peers{};
    peers.world = world;
    for (int r = 0; r < world; ++r) peers.dst[r] = reinterpret_cast<int32_t*>(peer_bases[r]) + dst_off_words;
    launch_dp_push(h->st, h->dp_block, L.stride, peers);
    CK(cudaGetLastError());
    return SCORE_OK;
}

// Optimizer half of the data-parallel step on the gathered blocks of all ranks (`gathered`: world blocks of
// score_dp_block_words(cap) words, rank order, device memory): dense gradients summed in rank order + dense Adam; the
// ranks' id lists merged into (id, rank) order (no sort: every list is already ascending), rows of one id added in
// rank order, fused row Adam - the same update on every replica, bit for bit.
// loss_out == NULL: enqueue only; otherwise *loss_out = the GLOBAL loss (sum of the ranks' data terms + the L2 term).
int score_dp_finish(ScoreHandle h, const void* gathered, int32_t world, int64_t cap, double* loss_out) {
    if (!h || !gathered || world < 1) return SCORE_ERR_ARG;
    if (!h->begun) return fail(h, SCORE_ERR_ARG, "score_dp_finish needs score_step_begin");
    if (cap <= 0 || cap % 1024) return fail(h, SCORE_ERR_ARG, "cap must be a positive multiple of 1024");
    CK(cudaSetDevice(h->device));
    const DpLayout L = dp_layout(h, gathered, world, cap);
    if ((int64_t)world * cap * (h->dm.d >> 2) >= ((int64_t)1 << 40)) return fail(h, SCORE_ERR_ARG, "gathered lists too large");
    launch_dp_dense_adam(h->st, L, h->P, h->M1, h->V1, h->G, h->flags, (int)h->n_dense, h->hyper_dev, h->alpha_hist, h->loss_glob);
    EmbUpdateArgs ea{};
    ea.d = h->dm.d;
    ea.emb = h->emb; ea.m = h->emb_m; ea.v = h->emb_v; ea.es = h->es; ea.last_step = h->last_step;
    ea.alpha_hist = h->alpha_hist; ea.hp = h->hyper_dev; ea.mode = 0;
    if (dp_sample_count(world, cap) > h->dp_sample_cap) {
        CK(cudaStreamSynchronize(h->st));
        if (h->dp_sample) cudaFree(h->dp_sample);
        h->dp_sample = nullptr; h->dp_sample_cap = 0;
        const int64_t want = 2 * dp_sample_count(world, cap);
        CK(cudaMalloc(&h->dp_sample, sizeof(int32_t) * want));
        h->dp_sample_cap = want;
    }
    launch_dp_apply(h->st, L, ea, h->dp_sample, h->err_flag);
    if (h->local_sorted) { h->last_sorted = h->sb.keys[h->sort_out]; h->last_sorted_n = h->dm.N; }   // this rank's share (stats)
    if (h->cfg.adam_mode == SCORE_ADAM_DENSE)
        launch_emb_dense_sweep(h->st, h->emb, h->emb_m, h->emb_v, h->last_step, h->dm.V, h->dm.d, h->es, h->hyper_dev);
    h->step += 1;
    h->beta1_power = h->beta1_power * 0.9f;
    h->beta2_power = h->beta2_power * 0.999f;
    h->begun = false;
    CK(cudaGetLastError());
    if (!loss_out) return SCORE_OK;   // asynchronous: the caller collects errors later with score_wait()
    CK(cudaMemcpyAsync(h->loss_glob_host, h->loss_glob, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    const int rc = finish_sync(h, nullptr);
    *loss_out = *h->loss_glob_host;
    return rc;
}

// Optimizer half of the split step: dense Adam on the (all-reduced) "dense_grad" buffer and the deterministic
// sort + segment-reduce + row Adam over an externally assembled (key, gradient row) list - the concatenation of
// every rank's positions for a replicated table, or the rows this rank owns for a sharded one.
// loss2[0] = this rank's loss (data part uses 1/global_batch) incl. the L2 term, loss2[1] = the L2 term alone.
int score_step_finish(ScoreHandle h, const int32_t* ext_keys, const float* ext_rows, int64_t n_ext, float* loss2) {
    if (!h) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (h->begun) {
        if (n_ext >= ((int64_t)1 << 31)) return fail(h, SCORE_ERR_ARG, "too many gradient rows for int32 positions");
        launch_dense_adam(h->st, h->P, h->M1, h->V1, h->G, h->flags, (int)h->n_dense, h->hyper_dev, h->alpha_hist);
        if (n_ext > 0) {
            int out;
            if (h->presorted_keys == ext_keys && h->presorted_n == n_ext) {   // score_shard_presort did it on the side stream
                out = h->presorted_out;
                CK(cudaStreamWaitEvent(h->st, h->ev_presort, 0));
            } else {
                int rc = ensure_ext_sort(h, n_ext);
                if (rc) return rc;
                out = launch_sort_pairs(h->st, h->sb_ext, ext_keys, n_ext, key_bits(h->dm.V));
                launch_emb_runs(h->st, h->sb_ext.keys[out], h->sb_ext.vals[out], n_ext, h->sb_ext.runs, h->sb_ext.runs_long, h->n_heads_dev, h->sb_ext.slotinfo);
            }
            h->presorted_keys = nullptr; h->presorted_n = 0;
            EmbUpdateArgs ea{};
            ea.skeys = h->sb_ext.keys[out]; ea.spos = h->sb_ext.vals[out]; ea.n = n_ext;
            ea.runs = h->sb_ext.runs; ea.runs_long = h->sb_ext.runs_long; ea.long_cap = emb_runs_long_cap(n_ext); ea.counters = h->n_heads_dev;
            ea.part = h->sb_ext.part; ea.slotinfo = h->sb_ext.slotinfo; ea.done = h->sb_ext.done;
            ea.grad_rows = ext_rows; ea.d = h->dm.d;
            ea.emb = h->emb; ea.m = h->emb_m; ea.v = h->emb_v; ea.es = h->es; ea.last_step = h->last_step;
            ea.alpha_hist = h->alpha_hist; ea.hp = h->hyper_dev; ea.mode = 0;
            h->last_sorted = ea.skeys; h->last_sorted_n = ea.n;
            launch_emb_update(h->st, ea);
        }
        if (h->cfg.adam_mode == SCORE_ADAM_DENSE)
            launch_emb_dense_sweep(h->st, h->emb, h->emb_m, h->emb_v, h->last_step, h->dm.V, h->dm.d, h->es, h->hyper_dev);
        h->step += 1;
        h->beta1_power = h->beta1_power * 0.9f;
        h->beta2_power = h->beta2_power * 0.999f;
        h->begun = false;
    }
    if (!loss2) return SCORE_OK;   // asynchronous: the caller collects errors / the loss later with score_wait()
    float l = 0.f;
    int rc = finish_sync(h, &l);
    loss2[0] = h->loss_host[0]; loss2[1] = h->loss_host[1];
    return rc;
}

}  // extern "C"

extern "C" {

int score_set_sample_offset(ScoreHandle h, int32_t first_global_sample) {
    if (!h || first_global_sample < 0) return SCORE_ERR_ARG;
    h->sample_base = first_global_sample;
    return SCORE_OK;
}

int score_stream(ScoreHandle h, void** cuda_stream) {
    if (!h || !cuda_stream) return SCORE_ERR_ARG;
    *cuda_stream = (void*)h->st;
    return SCORE_OK;
}

int64_t score_launch_count(ScoreHandle h) { return h ? g_launch_count - h->launches0 : 0; }

int score_enable_probes(ScoreHandle h, int on) {
    if (!h) return SCORE_ERR_ARG;
    const bool want = on != 0;
    for (int p = 0; p < PR_COUNT; ++p) { h->pr_ms[p] = 0; h->pr_n[p] = 0; }
    if (want == h->probes_on) return SCORE_OK;   // same state: only reset the accumulators
    h->probes_on = want;
    // graphs captured with / without the probe event nodes must be rebuilt
    for (auto& kv : h->graphs_train) cudaGraphExecDestroy(kv.second);
    h->graphs_train.clear();
    h->graph_kernels.clear();
    for (auto& kv : h->graphs_begin) cudaGraphExecDestroy(kv.second);
    h->graphs_begin.clear();
    h->graph_kernels_begin.clear();
    return SCORE_OK;
}

// out[2*p] = accumulated ms, out[2*p+1] = samples, for p in coatt_fwd, coatt_bwd, emb_update, sort, step,
// fwd_dense, bwd_dense, catchup
int score_probe_times(ScoreHandle h, double* out, int n) {
    if (!h || !out) return SCORE_ERR_ARG;
    for (int p = 0; p < PR_COUNT && 2 * p + 1 < n; ++p) { out[2 * p] = h->pr_ms[p]; out[2 * p + 1] = (double)h->pr_n[p]; }
    return PR_COUNT;
}

// bytes-level facts of the last step for roofline arithmetic: out = {N positions, live (non-zero key) positions, unique rows}
int score_last_step_stats(ScoreHandle h, int64_t* out3) {
    if (!h || !out3) return SCORE_ERR_ARG;
    CK(cudaSetDevice(h->device));
    const int64_t N = h->last_sorted_n;
    if (!h->last_sorted || N <= 0) return fail(h, SCORE_ERR_ARG, "no optimizer step has run yet");
    std::vector<int32_t> sk((size_t)N);
    CK(cudaStreamSynchronize(h->st));
    CK(cudaMemcpy(sk.data(), h->last_sorted, sizeof(int32_t) * N, cudaMemcpyDeviceToHost));
    int64_t live = 0, uniq = 0;
    for (int64_t i = 0; i < N; ++i) {
        if (sk[i] == 0) continue;
        ++live;
        if (i == 0 || sk[i - 1] != sk[i]) ++uniq;
    }
    out3[0] = N; out3[1] = live; out3[2] = uniq;
    return SCORE_OK;
}

}  // extern "C"
