// Evaluation metrics of train_score.py on the device.
//
// Replaces the arithmetic of eval() (train_score.py:144-163: sklearn log_loss + roc_auc_score over
// all predictions) and get_ranking_quality / getNDCG_at_K / getHR_at_K / getMRR
// (train_score.py:104-142: per group of 1 positive + 99 negatives, descending rank list via
// reversed(argsort(preds)), first occurrence of the positive item id).
//
// Tie rule: np.argsort's default introsort is unstable, so the reference's rank of a tied positive
// is implementation-defined.  This kernel uses the stable rule (ascending stable argsort, then
// reversed): among equal predictions the LATER candidate ranks first, i.e.
//     rank(j) = #{k : pred_k > pred_j} + #{k > j : pred_k == pred_j}.
// Tie-free groups match the reference exactly.
#include "kernels.h"

namespace score {

// one warp per group; out[g] = 0-based position of the first entry of the rank list whose item id
// equals the group's positive item id (candidate 0), or `group` if none (cannot happen: j = 0 matches)
__global__ void group_rank_kernel(const float* __restrict__ preds, const int32_t* __restrict__ iids, int64_t n_groups,
                                  int group, int32_t* __restrict__ pos_rank) {
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    const float* p = preds + g * group;
    const int32_t* id = iids + g * group;
    const int32_t target = id[0];
    int best = group;
    for (int j = lane; j < group; j += 32) {
        if (id[j] != target) continue;
        const float pj = p[j];
        int r = 0;
        for (int k = 0; k < group; ++k) r += (p[k] > pj) || (p[k] == pj && k > j);
        best = min(best, r);
    }
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(FULL_MASK, best, o));
    if (lane == 0) pos_rank[g] = best;
}

// per-group metric terms in double: ndcg@5, ndcg@10, hr@1, hr@5, hr@10, mrr  (train_score.py:104-120)
__global__ void rank_terms_kernel(const int32_t* __restrict__ pos_rank, int64_t n_groups, double* __restrict__ terms) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const int p = pos_rank[g];
    const double ndcg = log(2.0) / log((double)p + 2.0);
    double* t = terms + g * 6;
    t[0] = p < 5 ? ndcg : 0.0;
    t[1] = p < 10 ? ndcg : 0.0;
    t[2] = p < 1 ? 1.0 : 0.0;
    t[3] = p < 5 ? 1.0 : 0.0;
    t[4] = p < 10 ? 1.0 : 0.0;
    t[5] = 1.0 / ((double)p + 1.0);
}

// sklearn log_loss term: probabilities [1-p, p] clipped to [eps, 1-eps], eps = float64 machine epsilon
__global__ void logloss_terms_kernel(const float* __restrict__ preds, const int32_t* __restrict__ labels, int64_t n,
                                     double* __restrict__ terms) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double eps = 2.220446049250313e-16;
    double p = (double)preds[i];
    double q = labels[i] ? p : 1.0 - p;
    q = fmin(fmax(q, eps), 1.0 - eps);
    terms[i] = -log(q);
}

// column sums of a [rows, cols] double matrix, one block per column, fixed-shape tree
__global__ void colsum_f64_kernel(const double* __restrict__ x, int64_t rows, int cols, double* __restrict__ out) {
    __shared__ double red[256];
    const int c = blockIdx.x;
    double s = 0.0;
    for (int64_t r = threadIdx.x; r < rows; r += 256) s += x[r * cols + c];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = red[0];
}

// float -> uint32 whose unsigned order is the float order
__global__ void auc_keys_kernel(const float* __restrict__ preds, int64_t n, int32_t* __restrict__ keys) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t u = __float_as_uint(preds[i]);
    u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;
    keys[i] = (int32_t)u;
}

// Mann-Whitney with mid-ranks (== area under sklearn's ROC curve): every head of a run of equal
// predictions adds (#positives in run) * (mean 1-based rank of the run)
__global__ void auc_terms_kernel(const int32_t* __restrict__ skeys, const int32_t* __restrict__ spos,
                                 const int32_t* __restrict__ labels, int64_t n, double* __restrict__ terms) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double t0 = 0.0, t1 = 0.0;
    if (i == 0 || skeys[i - 1] != skeys[i]) {
        int64_t j = i, npos = 0;
        for (; j < n && skeys[j] == skeys[i]; ++j) npos += labels[spos[j]] != 0;
        const double len = (double)(j - i);
        t0 = (double)npos * ((double)i + (len + 1.0) * 0.5);
        t1 = (double)npos;
    }
    terms[i * 2] = t0;
    terms[i * 2 + 1] = t1;
}

// out9 = logloss, auc, ndcg5, ndcg10, hr1, hr5, hr10, mrr, 0 ; all buffers device, scratch caller-provided
cudaError_t compute_eval_metrics(cudaStream_t st, const float* preds, const int32_t* iids, const int32_t* labels,
                                 int64_t n, int group, SortBufs& sb, int32_t* keys, int32_t* pos_rank, double* terms,
                                 double* sums /* [16] */) {
    const int64_t n_groups = n / group;
    const unsigned nb = (unsigned)((n + 255) / 256);
    // ranking metrics
    group_rank_kernel<<<(unsigned)((n_groups * 32 + 127) / 128), 128, 0, st>>>(preds, iids, n_groups, group, pos_rank);
    rank_terms_kernel<<<(unsigned)((n_groups + 255) / 256), 256, 0, st>>>(pos_rank, n_groups, terms);
    colsum_f64_kernel<<<6, 256, 0, st>>>(terms, n_groups, 6, sums + 2);
    // log loss
    logloss_terms_kernel<<<nb, 256, 0, st>>>(preds, labels, n, terms);
    colsum_f64_kernel<<<1, 256, 0, st>>>(terms, n, 1, sums + 0);
    // AUC
    auc_keys_kernel<<<nb, 256, 0, st>>>(preds, n, keys);
    g_launch_count += 6;
    int out = launch_sort_pairs(st, sb, keys, n, 32);
    auc_terms_kernel<<<nb, 256, 0, st>>>(sb.keys[out], sb.vals[out], labels, n, terms);
    colsum_f64_kernel<<<2, 256, 0, st>>>(terms, n, 2, sums + 8);
    g_launch_count += 2;
    return cudaGetLastError();
}

}  // namespace score
