// On-GPU graph store + neighbor sampler: the device replacement of GraphHandler / GraphLoader
// (code/score/graph_loader.py:94-277, 340-385).  SURVEY.md section 8f-1.
//
// The reference keeps one Mongo document per node - {'1hop': [S lists], '2hop': [S lists <= 100 ids],
// 'degrees': [S lists]} (graph_storage.py:153-245) - and nine Python worker processes turn (uid, iid) targets into
// nested lists.  Here the whole graph is two CSR arrays in HBM, indexed by (node, slice):
//     hop1_off[node*S + s] .. hop1_off[node*S + s + 1]  ->  hop1_ids      (interaction partners in slice s)
//     hop2_off[node*S + s] .. hop2_off[node*S + s + 1]  ->  hop2_ids, hop2_deg
// plus the side-feature tables (item -> [cid, sid, bid], user -> [aid, gid], feateng_tmall.py:118-133), and one
// kernel writes the six id tensors of a batch straight into device memory in the layout score_train_step consumes
// with on_device = 1.  Per (entity, slice t, neighbor k), as GraphHandler.gen_*_neighbor_{rs,is} do:
//   1-hop   n = len(list):  n == 0 -> dummy node (all fields 0);  else list[k mod n]
//           (first K if longer, cyclic repeat if shorter: graph_loader.py:181-185)
//   2-hop   n == 0 -> dummy;  'rs': list[floor(u*n)], u uniform (np.random.choice(list, K), :203);
//           'is': inverse CDF of softmax(1 / (degree - 1)) at u (:112-114)
//   every id is expanded to [id] + feat[id]  (:186-191);  slices t >= pred_time - start_time copy the last live
//   slice (:254-256);  the user side is computed once per target user and replicated over its 1 + neg samples
//   (:360-364);  label = 1 for the first sample of a group, length = pred_time - start_time (:378-382).
// The uniforms are Philox4x32-10 keyed by (seed, draw id, side, entity, slice, k), so a batch is reproducible and
// the CPU restatement (oracle/loader_ref.py) can be handed the same draws; NumPy's generator cannot be reproduced.
#include <string>
#include <vector>

#include "../../include/score_b200.h"
#include "kernels.h"

namespace score {

struct GraphDev {
    int n_user = 0, n_item = 0, S = 0, uf = 1, fi = 1;
    int64_t n_nodes = 0;          // n_user + n_item + 1 (row 0 unused)
    int64_t* hop1_off = nullptr; int32_t* hop1_ids = nullptr;
    int64_t* hop2_off = nullptr; int32_t* hop2_ids = nullptr; int32_t* hop2_deg = nullptr;
    int32_t* user_feat = nullptr;   // [(n_user + 1) * (uf - 1)]
    int32_t* item_feat = nullptr;   // [(n_item + 1) * (fi - 1)]
};

struct SampleArgs {
    GraphDev g;
    const int32_t* uids; const int32_t* iids;   // device: [n_groups], [B]
    int B, grp, T, K, L, start_time, mode;
    uint32_t seed_lo, seed_hi, draw_id;
    int32_t *u1, *u2, *i1, *i2, *tu, *ti, *label, *length;
    int32_t* err_flag;
};

// [id] + feat[id] for a user-typed (is_user) or item-typed node; id 0 -> zeros
__device__ __forceinline__ void write_node(const GraphDev& g, int32_t id, bool is_user, int32_t* dst) {
    const int f = is_user ? g.uf : g.fi;
    dst[0] = id;
    if (f == 1) return;
    if (id == 0) {
        for (int j = 1; j < f; ++j) dst[j] = 0;
        return;
    }
    const int32_t* src = is_user ? g.user_feat + (int64_t)id * (g.uf - 1)
                                 : g.item_feat + (int64_t)(id - g.n_user) * (g.fi - 1);
    for (int j = 1; j < f; ++j) dst[j] = src[j - 1];
}

__device__ __forceinline__ int32_t pick_1hop(const GraphDev& g, int32_t node, int s, int k) {
    const int64_t o0 = g.hop1_off[(int64_t)node * g.S + s], o1 = g.hop1_off[(int64_t)node * g.S + s + 1];
    const int n = (int)(o1 - o0);
    return n == 0 ? 0 : g.hop1_ids[o0 + (k % n)];
}

__device__ __forceinline__ int32_t pick_2hop(const GraphDev& g, int32_t node, int s, float u, int mode) {
    const int64_t o0 = g.hop2_off[(int64_t)node * g.S + s], o1 = g.hop2_off[(int64_t)node * g.S + s + 1];
    const int n = (int)(o1 - o0);
    if (n == 0) return 0;
    if (mode == 0) {   // 'rs'
        int j = (int)(u * (float)n);
        if (j >= n) j = n - 1;
        return g.hop2_ids[o0 + j];
    }
    // 'is': p_i = softmax(1 / (deg_i - 1)); NumPy draws with cdf.searchsorted(u, side='right') in float64
    double mx = -1e300;
    for (int i = 0; i < n; ++i) { const double x = 1.0 / ((double)g.hop2_deg[o0 + i] - 1.0); mx = x > mx ? x : mx; }
    double tot = 0.0;
    for (int i = 0; i < n; ++i) tot += exp(1.0 / ((double)g.hop2_deg[o0 + i] - 1.0) - mx);
    const double target = (double)u * tot;
    double c = 0.0;
    int j = n - 1;
    for (int i = 0; i < n; ++i) {
        c += exp(1.0 / ((double)g.hop2_deg[o0 + i] - 1.0) - mx);
        if (c > target) { j = i; break; }
    }
    return g.hop2_ids[o0 + j];
}

// one thread per (entity, t, k); entities: the n_groups target users first, then the B target items
__global__ void sample_batch_kernel(SampleArgs a) {
    const GraphDev& g = a.g;
    const int n_groups = (a.B + a.grp - 1) / a.grp;
    const int64_t per = (int64_t)a.T * a.K;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < a.B) {   // targets, label, length
        const int b = (int)idx;
        const int32_t uid = a.uids[b / a.grp], iid = a.iids[b];
        const bool ok = uid >= 1 && uid <= g.n_user && iid > g.n_user && iid <= g.n_user + g.n_item;
        if (!ok) atomicExch(a.err_flag, 1);
        write_node(g, ok ? uid : 0, true, a.tu + (int64_t)b * g.uf);
        write_node(g, ok ? iid : 0, false, a.ti + (int64_t)b * g.fi);
        a.label[b] = (b % a.grp == 0) ? 1 : 0;
        a.length[b] = a.L;
    }
    if (idx >= (int64_t)(n_groups + a.B) * per) return;
    const int64_t ent = idx / per;
    const int rem = (int)(idx - ent * per);
    const int t = rem / a.K, k = rem - t * a.K;
    const int ts = (t < a.L ? t : a.L - 1);            // tail slices copy the last live one
    const int s = a.start_time + ts;
    const bool user_side = ent < n_groups;
    const int32_t node = user_side ? a.uids[ent] : a.iids[ent - n_groups];
    const bool ok = user_side ? (node >= 1 && node <= g.n_user) : (node > g.n_user && node <= g.n_user + g.n_item);
    int32_t id1 = 0, id2 = 0;
    if (ok && s < g.S) {
        id1 = pick_1hop(g, node, s, k);
        const uint64_t key = ((uint64_t)ent * (uint64_t)a.T + (uint64_t)ts) * (uint64_t)a.K + (uint64_t)k;
        const float u = philox_uniform(a.seed_lo, a.seed_hi, user_side ? 1u : 2u, a.draw_id, key);
        id2 = pick_2hop(g, node, s, u, a.mode);
    }
    if (user_side) {
        // user_1hop holds item-typed nodes, user_2hop user-typed nodes; replicate over the group's samples
        const int b0 = (int)ent * a.grp;
        for (int j = 0; j < a.grp && b0 + j < a.B; ++j) {
            const int64_t o = ((int64_t)(b0 + j) * a.T + t) * a.K + k;
            write_node(g, id1, false, a.u1 + o * g.fi);
            write_node(g, id2, true, a.u2 + o * g.uf);
        }
    } else {
        const int64_t o = ((ent - n_groups) * a.T + t) * a.K + k;
        write_node(g, id1, true, a.i1 + o * g.uf);     // item_1hop: user-typed
        write_node(g, id2, false, a.i2 + o * g.fi);    // item_2hop: item-typed
    }
}

}  // namespace score

using namespace score;

struct ScoreGraph {
    GraphDev g;
    int device = 0;
    cudaStream_t st = nullptr;
    std::string err;
    // output buffers of the last sample call (grown on demand)
    int32_t *u1 = nullptr, *u2 = nullptr, *i1 = nullptr, *i2 = nullptr, *tu = nullptr, *ti = nullptr, *label = nullptr,
            *length = nullptr, *uids = nullptr, *iids = nullptr, *err_flag = nullptr;
    int64_t cap_hist = 0; int cap_B = 0;
};

static std::string g_graph_error;
#define GCK(x)                                                                                         \
    do {                                                                                               \
        cudaError_t e_ = (x);                                                                          \
        if (e_ != cudaSuccess) { h->err = std::string(#x) + ": " + cudaGetErrorString(e_); return SCORE_ERR_CUDA; } \
    } while (0)

template <class T>
static cudaError_t upload(T** dst, const T* src, size_t n) {
    *dst = nullptr;
    if (n == 0) n = 1;
    cudaError_t e = cudaMalloc(dst, n * sizeof(T));
    if (e != cudaSuccess) return e;
    if (src) return cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice);
    return cudaMemset(*dst, 0, n * sizeof(T));
}

extern "C" {

const char* score_graph_last_error(ScoreGraphHandle h) { return h ? h->err.c_str() : g_graph_error.c_str(); }

// The sampler indexes the side-feature tables with the neighbor ids it picks ([id] + feat[id], graph_loader.py:186-191,
// where an id without a feature entry is a KeyError): the graph is bipartite by construction (graph_storage.py:127-246),
// so a user's 1-hop entries and an item's 2-hop entries are item ids, a user's 2-hop entries and an item's 1-hop entries
// are user ids; 0 is the dummy.  Checked once here, on the host arrays, before anything is uploaded.
static bool validate_csr(const ScoreGraphDesc* d, const int64_t* off, const int32_t* ids, bool hop1, std::string* msg) {
    const int64_t S = d->n_slices, n_lists = ((int64_t)d->n_user + d->n_item + 1) * S;
    const char* what = hop1 ? "hop1" : "hop2";
    if (off[0] != 0) { *msg = std::string(what) + "_off[0] must be 0"; return false; }
    for (int64_t i = 0; i < n_lists; ++i)
        if (off[i + 1] < off[i]) { *msg = std::string(what) + "_off is not ascending"; return false; }
    if (off[n_lists] > 0 && !ids) { *msg = std::string(what) + "_ids is NULL"; return false; }
    for (int64_t node = 1; node <= (int64_t)d->n_user + d->n_item; ++node) {
        const bool node_is_user = node <= d->n_user;
        const bool want_user = node_is_user != hop1;      // 1-hop neighbors are of the other type, 2-hop of the own type
        const int32_t lo = want_user ? 1 : d->n_user + 1, hi = want_user ? d->n_user : d->n_user + d->n_item;
        for (int64_t i = off[node * S]; i < off[(node + 1) * S]; ++i) {
            const int32_t v = ids[i];
            if (v != 0 && (v < lo || v > hi)) {
                *msg = std::string(what) + " list of node " + std::to_string(node) + " holds id " + std::to_string(v) + ", expected 0 or " +
                       std::to_string(lo) + ".." + std::to_string(hi) + (want_user ? " (a user id)" : " (an item id)");
                return false;
            }
        }
    }
    return true;
}

int score_graph_create(const ScoreGraphDesc* d, int device, ScoreGraphHandle* out) {
    if (!d || !out) { g_graph_error = "null argument"; return SCORE_ERR_ARG; }
    if (d->n_user <= 0 || d->n_item <= 0 || d->n_slices <= 0 || d->user_fnum < 1 || d->item_fnum < 1 || !d->hop1_off ||
        !d->hop2_off) { g_graph_error = "bad graph description"; return SCORE_ERR_ARG; }
    if (!validate_csr(d, d->hop1_off, d->hop1_ids, true, &g_graph_error) || !validate_csr(d, d->hop2_off, d->hop2_ids, false, &g_graph_error))
        return SCORE_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        g_graph_error = "no CUDA device: the graph store has no CPU fallback";
        cudaGetLastError();
        return SCORE_ERR_CUDA;
    }
    ScoreGraph* h = new ScoreGraph();
    h->device = device;
    auto die = [&](const std::string& m, int code) { g_graph_error = m; score_graph_destroy(h); return code; };
    if (cudaSetDevice(device) != cudaSuccess) return die("cudaSetDevice failed", SCORE_ERR_CUDA);
    GraphDev& g = h->g;
    g.n_user = d->n_user; g.n_item = d->n_item; g.S = d->n_slices; g.uf = d->user_fnum; g.fi = d->item_fnum;
    g.n_nodes = (int64_t)d->n_user + d->n_item + 1;
    const size_t n_off = (size_t)g.n_nodes * g.S + 1;
    const int64_t n1 = d->hop1_off[n_off - 1], n2 = d->hop2_off[n_off - 1];
    if (n1 < 0 || n2 < 0) return die("negative CSR size", SCORE_ERR_ARG);
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = upload(&g.hop1_off, d->hop1_off, n_off);
    if (e == cudaSuccess) e = upload(&g.hop1_ids, d->hop1_ids, (size_t)n1);
    if (e == cudaSuccess) e = upload(&g.hop2_off, d->hop2_off, n_off);
    if (e == cudaSuccess) e = upload(&g.hop2_ids, d->hop2_ids, (size_t)n2);
    if (e == cudaSuccess) e = upload(&g.hop2_deg, d->hop2_deg, (size_t)n2);   // NULL -> zeros ('rs' mode only)
    if (e == cudaSuccess) e = upload(&g.user_feat, d->user_feat, (size_t)(d->n_user + 1) * (g.uf - 1));
    if (e == cudaSuccess) e = upload(&g.item_feat, d->item_feat, (size_t)(d->n_item + 1) * (g.fi - 1));
    if (e == cudaSuccess) e = cudaMalloc(&h->err_flag, sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemset(h->err_flag, 0, sizeof(int32_t));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
    if (e != cudaSuccess) return die(std::string("graph upload failed: ") + cudaGetErrorString(e), SCORE_ERR_CUDA);
    *out = h;
    return SCORE_OK;
}

int score_graph_destroy(ScoreGraphHandle h) {
    if (!h) return SCORE_OK;
    cudaSetDevice(h->device);
    GraphDev& g = h->g;
    for (void* p : {(void*)g.hop1_off, (void*)g.hop1_ids, (void*)g.hop2_off, (void*)g.hop2_ids, (void*)g.hop2_deg,
                    (void*)g.user_feat, (void*)g.item_feat, (void*)h->u1, (void*)h->u2, (void*)h->i1, (void*)h->i2,
                    (void*)h->tu, (void*)h->ti, (void*)h->label, (void*)h->length, (void*)h->uids, (void*)h->iids,
                    (void*)h->err_flag})
        if (p) cudaFree(p);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
    return SCORE_OK;
}

int score_graph_sample(ScoreGraphHandle h, const int32_t* uids, const int32_t* iids, int32_t batch_size,
                       int32_t group, int32_t start_time, int32_t pred_time, int32_t max_time_len,
                       int32_t obj_per_time_slice, int32_t mode, uint64_t seed, uint32_t draw_id, void* cuda_stream,
                       ScoreBatch* out) {
    if (!h) return SCORE_ERR_ARG;
    auto bad = [&](const char* m) { h->err = m; return SCORE_ERR_ARG; };
    if (!uids || !iids || !out) return bad("null argument");
    const int B = batch_size, T = max_time_len, K = obj_per_time_slice, L = pred_time - start_time;
    if (B <= 0 || group <= 0 || T <= 0 || K <= 0) return bad("batch_size, group, max_time_len, obj_per_time_slice must be positive");
    if (L < 1 || L > T) return bad("pred_time - start_time must be in [1, max_time_len]");
    if (start_time < 0 || pred_time > h->g.S) return bad("time slices outside the stored graph");
    if (mode != 0 && mode != 1) return bad("mode must be 0 ('rs') or 1 ('is')");
    GCK(cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->st;
    const GraphDev& g = h->g;
    const int n_groups = (B + group - 1) / group;
    const int64_t hist = (int64_t)B * T * K;
    if (hist > h->cap_hist || B > h->cap_B) {
        GCK(cudaStreamSynchronize(st));
        for (void* p : {(void*)h->u1, (void*)h->u2, (void*)h->i1, (void*)h->i2, (void*)h->tu, (void*)h->ti,
                        (void*)h->label, (void*)h->length, (void*)h->uids, (void*)h->iids})
            if (p) cudaFree(p);
        GCK(cudaMalloc(&h->u1, sizeof(int32_t) * hist * g.fi));
        GCK(cudaMalloc(&h->u2, sizeof(int32_t) * hist * g.uf));
        GCK(cudaMalloc(&h->i1, sizeof(int32_t) * hist * g.uf));
        GCK(cudaMalloc(&h->i2, sizeof(int32_t) * hist * g.fi));
        GCK(cudaMalloc(&h->tu, sizeof(int32_t) * B * g.uf));
        GCK(cudaMalloc(&h->ti, sizeof(int32_t) * B * g.fi));
        GCK(cudaMalloc(&h->label, sizeof(int32_t) * B));
        GCK(cudaMalloc(&h->length, sizeof(int32_t) * B));
        GCK(cudaMalloc(&h->uids, sizeof(int32_t) * B));
        GCK(cudaMalloc(&h->iids, sizeof(int32_t) * B));
        h->cap_hist = hist; h->cap_B = B;
    }
    GCK(cudaMemcpyAsync(h->uids, uids, sizeof(int32_t) * n_groups, cudaMemcpyDefault, st));
    GCK(cudaMemcpyAsync(h->iids, iids, sizeof(int32_t) * B, cudaMemcpyDefault, st));
    SampleArgs a{};
    a.g = g; a.uids = h->uids; a.iids = h->iids; a.B = B; a.grp = group; a.T = T; a.K = K; a.L = L;
    a.start_time = start_time; a.mode = mode;
    a.seed_lo = (uint32_t)seed; a.seed_hi = (uint32_t)(seed >> 32); a.draw_id = draw_id;
    a.u1 = h->u1; a.u2 = h->u2; a.i1 = h->i1; a.i2 = h->i2; a.tu = h->tu; a.ti = h->ti; a.label = h->label;
    a.length = h->length; a.err_flag = h->err_flag;
    const int64_t threads = (int64_t)(n_groups + B) * T * K;
    sample_batch_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(a);
    ++g_launch_count;
    GCK(cudaGetLastError());
    out->user_1hop = h->u1; out->user_2hop = h->u2; out->item_1hop = h->i1; out->item_2hop = h->i2;
    out->target_user = h->tu; out->target_item = h->ti; out->label = h->label; out->length = h->length;
    out->batch_size = B; out->on_device = 1;
    return SCORE_OK;
}

/* wait for the sampler's stream and report ids outside the stored graph */
int score_graph_sync(ScoreGraphHandle h, void* cuda_stream) {
    if (!h) return SCORE_ERR_ARG;
    GCK(cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->st;
    int32_t flag = 0;
    GCK(cudaMemcpyAsync(&flag, h->err_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    GCK(cudaStreamSynchronize(st));
    if (flag) {
        cudaMemsetAsync(h->err_flag, 0, sizeof(int32_t), st);
        h->err = "a target uid / iid is outside the stored graph";
        return SCORE_ERR_ID_RANGE;
    }
    return SCORE_OK;
}

/* device -> host copy helper for the ids of a sampled batch (tests, debugging) */
int score_copy_to_host(void* dst_host, const void* src_device, size_t bytes) {
    return cudaMemcpy(dst_host, src_device, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? SCORE_OK : SCORE_ERR_CUDA;
}

}  // extern "C"
