/*
 * score_b200.h - C ABI of the B200-native SCoRe training / scoring hot path.
 *
 * This is the drop-in boundary for ONE path of qinjr/SCoRe: the model object that
 * code/score/train_score.py drives (construct :171, model.train :227, model.eval :153,
 * model.save :252, model.restore :80) and whose TensorFlow graph is built by
 * code/score/score.py (class SCORE :188-224 and the ablations :227-369).  Everything behind
 * these entry points is hand-written CUDA for sm_100a; there is no CPU fallback: every call
 * fails with SCORE_ERR_CUDA when no device / kernel image is available.
 *
 * Conventions
 *  - plain C types only; a handle owns every parameter, optimizer slot and workspace on ONE
 *    device (the reference keeps them in the tf.Session; one live model per process there);
 *  - int return: 0 = ok, otherwise a SCORE_ERR_* code; score_last_error() gives the text;
 *  - id tensors are int32, row-major, in the loader's layout (graph_loader.py:383):
 *      user_1hop [B,T,K,item_fnum]  user_2hop [B,T,K,user_fnum]
 *      item_1hop [B,T,K,user_fnum]  item_2hop [B,T,K,item_fnum]
 *      target_user [B,user_fnum]    target_item [B,item_fnum]   label [B]   length [B]
 *    every id indexes the single shared table (0 = dummy node, zero vector, no gradient);
 *  - calls on one handle are not re-entrant (train_score.py is single-threaded).
 */
#ifndef SCORE_B200_H
#define SCORE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ScoreModel* ScoreHandle;

enum {
    SCORE_OK = 0,
    SCORE_ERR_ARG = 1,      /* bad argument / shape (reference: TF raises from sess.run)          */
    SCORE_ERR_CUDA = 2,     /* CUDA runtime error, no device, or no sm_100a kernel image           */
    SCORE_ERR_ID_RANGE = 3, /* an id outside [0, feature_size) (reference: embedding_lookup raises) */
    SCORE_ERR_IO = 4,       /* save / restore file error                                           */
    SCORE_ERR_NAME = 5      /* unknown tensor name                                                 */
};

/* model_type: which class of score.py the handle mirrors (train_score.py:170-182); RRN is the slice baseline of
 * code/slice_models/slice_model.py:155-173 (same constructor, batch and train / eval surface: 1-hop sum pooling,
 * two GRUs on per-side widths item_fnum*d / user_fnum*d, final states into the prediction MLP). */
enum { SCORE_MODEL_SCORE = 0, SCORE_MODEL_RIA = 1, SCORE_MODEL_RCA = 2,
       SCORE_MODEL_SCORE_USER = 3, SCORE_MODEL_SCORE_ITEM = 4, SCORE_MODEL_RRN = 5 };

/* Embedding optimizer mode.  The reference's emb_mtx gradient is dense (score.py:45-47 route the
 * lookup through a dense multiply), so tf.train.AdamOptimizer moves EVERY row EVERY step.
 *  DENSE  : literal - sparse row update for touched rows + a full-table zero-gradient sweep.
 *  LAZY   : exact same results as DENSE; a row's skipped zero-gradient steps are replayed
 *           (same fp32 op sequence) the next time the row is gathered, read back or saved.
 *  SPARSE : touched rows only (not step-for-step identical to the reference; perf mode).     */
enum { SCORE_ADAM_DENSE = 0, SCORE_ADAM_LAZY = 1, SCORE_ADAM_SPARSE = 2 };

/* Mirrors SCOREBASE.__init__(feature_size, eb_dim, hidden_size, max_time_len,
 * obj_per_time_slice, user_fnum, item_fnum)  (score.py:12-13), plus build options. */
typedef struct ScoreConfig {
    int64_t feature_size;       /* V: rows of emb_mtx                                  */
    int32_t eb_dim;             /* d (multiple of 4)                                   */
    int32_t hidden_size;        /* H                                                   */
    int32_t max_time_len;       /* T                                                   */
    int32_t obj_per_time_slice; /* K (<= 32)                                           */
    int32_t user_fnum;
    int32_t item_fnum;
    int32_t model_type;         /* SCORE_MODEL_*                                       */
    int32_t adam_mode;          /* SCORE_ADAM_*                                        */
    int32_t max_batch;          /* workspace capacity hint (grows on demand)           */
    uint64_t seed;              /* weight-init and dropout RNG seed                    */
    int32_t init_weights;       /* 1: draw TF-default initial weights on the device    */
    int32_t use_graph;          /* 1: replay the step as a CUDA graph (per batch size) */
} ScoreConfig;

/* Host id arrays of one batch, in batch_data order (score.py:103-110). */
typedef struct ScoreBatch {
    const int32_t* user_1hop;
    const int32_t* user_2hop;
    const int32_t* item_1hop;
    const int32_t* item_2hop;
    const int32_t* target_user;
    const int32_t* target_item;
    const int32_t* label;
    const int32_t* length;
    int32_t batch_size;         /* B (dynamic: the last batch is short, graph_loader.py:321-324) */
    int32_t on_device;          /* 0: host pointers (copied H2D inside the call); 1: device pointers */
} ScoreBatch;

/* replaces SCORE(...) construction + sess.run(global_variables_initializer) (train_score.py:171,188) */
int score_create(const ScoreConfig* cfg, int device, ScoreHandle* out);
int score_destroy(ScoreHandle h);
const char* score_last_error(ScoreHandle h); /* h may be NULL: error of the last failed create */

/* replaces SCOREBASE.train (score.py:101-116): one forward + backward + Adam update.
 * *loss_out is the pre-update loss incl. the L2 term (what sess.run([loss, train_step]) returns).
 * keep_prob is 0.8 in the reference (score.py:113). */
int score_train_step(ScoreHandle h, const ScoreBatch* batch, float lr, float reg_lambda,
                     float keep_prob, float* loss_out);
/* asynchronous variant: enqueue only; score_wait() returns the loss of the last enqueued step. */
int score_train_step_async(ScoreHandle h, const ScoreBatch* batch, float lr, float reg_lambda,
                           float keep_prob);
int score_wait(ScoreHandle h, float* loss_out);

/* replaces SCOREBASE.eval (score.py:118-133): keep_prob = 1; preds_out[B] = y_pred,
 * *loss_out = log-loss + L2 term. */
int score_eval(ScoreHandle h, const ScoreBatch* batch, float reg_lambda, float* preds_out,
               float* loss_out);

/* Parity / inspection: forward + backward WITHOUT the optimizer update.  Afterwards every
 * intermediate and gradient can be read with score_get_buffer().  Dense gradients are under
 * "grad/<tf var name>"; the embedding gradient row set under "emb_grad/rows" (int32, ascending)
 * and "emb_grad/values" ([U,d]). */
int score_forward_backward(ScoreHandle h, const ScoreBatch* batch, float reg_lambda,
                           float keep_prob, float* loss_out);
/* size query: data == NULL -> *count = number of elements.  dtype: 0 = float32, 1 = int32. */
int score_get_buffer(ScoreHandle h, const char* name, void* data, size_t capacity_bytes,
                     size_t* count, int* dtype);

/* Parameters and Adam slots by TF variable name ("emb_mtx", "dense/kernel", ...,
 * "<var>/Adam", "<var>/Adam_1", "beta1_power", "beta2_power").
 * Replaces reading / assigning tf variables; used for checkpoint interop and parity.
 * Setting a beta power also recovers the optimizer step from it (0.9^(step+1); 0.999^(step+1) once
 * beta1_power has gone denormal); "step" (one float, not a TF variable) sets / reads it exactly.
 * Either way the optimizer state present at that moment counts as current at that step. */
int score_tensor_count(ScoreHandle h);
int score_tensor_info(ScoreHandle h, int index, char* name, size_t name_cap, int64_t* rows, int64_t* cols);
int score_get_tensor(ScoreHandle h, const char* name, float* data, size_t count);
int score_set_tensor(ScoreHandle h, const char* name, const float* data, size_t count);
/* row-range access to emb_mtx and its slots for tables too large for one host buffer */
int score_get_rows(ScoreHandle h, const char* name, int64_t row0, int64_t nrows, float* data);
int score_set_rows(ScoreHandle h, const char* name, int64_t row0, int64_t nrows, const float* data);

/* replaces SCOREBASE.save / restore (score.py:135-142): all variables incl. Adam slots. */
int score_save(ScoreHandle h, const char* path);
int score_restore(ScoreHandle h, const char* path);

/* replaces the arithmetic of eval() + get_ranking_quality (train_score.py:122-163) for N
 * predictions in groups of `group` (100): out9 = logloss, auc, ndcg@5, ndcg@10, hr@1, hr@5,
 * hr@10, mrr, <unused 0>.  Host pointers. */
int score_eval_metrics(ScoreHandle h, const float* preds, const int32_t* target_iids,
                       const int32_t* labels, int64_t n, int32_t group, double* out9);

/* Multi-GPU split step (one process per GPU; the collectives themselves are issued by the host side with
 * torch.distributed / NCCL on score_stream(), between these calls).  There is no reference counterpart: the
 * reference is single-device; the split keeps SCOREBASE.train's arithmetic.
 *   data-parallel, replicated table:  score_step_begin -> score_dp_local_count (early, from the sort branch) ->
 *       counts exchanged -> score_dp_pack(cap) -> ONE all-gather of the packed blocks (dense gradient + one embedding-
 *       gradient row per unique id, ascending) -> score_dp_finish  (rank-ordered sums, merge instead of sort; the
 *       identical deterministic update on every replica)
 *   row-sharded table (owner = id % world):  score_prepare_batch -> all-to-all ids -> score_gather_rows on the
 *       owners -> all-to-all rows -> score_step_begin(staged table) -> all-reduce "dense_grad", all-to-all
 *       "grad_rows" to the owners -> score_step_finish(owned keys, owned rows).                              */
int score_prepare_batch(ScoreHandle h, const ScoreBatch* batch);
int score_device_buffer(ScoreHandle h, const char* name, void** dev_ptr, size_t* count);
int score_gather_rows(ScoreHandle h, const int32_t* idx_dev, int64_t n, float* out_dev);
int score_step_begin(ScoreHandle h, const ScoreBatch* batch, float lr, float reg_lambda, float keep_prob,
                     int32_t global_batch, int32_t train, const float* staged_table, const int32_t* staged_keys);
/* data-parallel only (see above).  A block of one rank is score_dp_block_words(cap) 4-byte words:
 *   [0,128) header {unique-row count, step sequence number, loss, L2 part}; [128, +n_dense) dense gradient;
 *   then cap ids (int32, ascending, zero-padded) and cap rows of eb_dim floats.  cap: multiple of 1024, the same on
 *   every rank, >= every rank's count.  score_dp_finish: loss_out == NULL enqueues only; otherwise *loss_out = the
 *   global loss (sum of the ranks' data terms, each scaled by 1/global_batch, + the L2 term). */
int score_dp_local_count(ScoreHandle h, int32_t* count_out);
int64_t score_dp_block_words(ScoreHandle h, int64_t cap);
int score_dp_pack(ScoreHandle h, int64_t cap, void** block_dev, int64_t* block_words);
int score_dp_finish(ScoreHandle h, const void* gathered_blocks_dev, int32_t world, int64_t cap, double* loss_out);
/* opt-in peer-memory exchange instead of the all-gather: store the packed block at word offset dst_off_words of every
 * replica's gathered buffer (peer_bases[r]: that buffer's address as mapped into this process); the caller runs a
 * barrier across the replicas on score_stream() before score_dp_finish. */
int score_dp_push(ScoreHandle h, int64_t cap, const uint64_t* peer_bases, int32_t world, int64_t dst_off_words);
/* loss2 == NULL: enqueue only (no host synchronisation); otherwise loss2[0] = this rank's loss incl. the L2 term
 * (data term scaled by 1/global_batch), loss2[1] = the L2 term alone. */
int score_step_finish(ScoreHandle h, const int32_t* ext_keys, const float* ext_rows, int64_t n_ext, float* loss2);

/* Row-sharded table, device side of the exchange (no reference counterpart).  After
 * score_prepare_batch: group this rank's positions by owner = id % world (position order inside a group, so the order in
 * which an owner receives and adds gradient rows is fixed).  All pointers are DEVICE memory owned by the handle, valid
 * until the next score_shard_plan; nothing is copied to the host (the caller all-gathers the count matrix):
 *   counts     [world + 1]          positions per owner; the last entry counts the dummy positions (id 0)
 *   send_rows  [n_valid]            owner-local row numbers (id / world + 1) in send order, n_valid = N - counts[world]
 *   staged     [(1 + N) * eb_dim]   row 0 unused; the rows the owners return land at rows 1.. in send order
 *   mini_keys  [N]                  position -> row of `staged` (0: dummy) - pass both to score_step_begin
 *   grad_send  [N * eb_dim]         score_shard_pack_grads: this rank's gradient rows in send order
 * score_shard_presort: sort the owner-side key list (the rows this rank serves, known before forward / backward) on the
 * side stream; score_step_finish then takes the same pointer and skips its own sort. */
typedef struct ScoreShardPlan {
    int32_t* counts; int32_t* send_rows; float* staged; int32_t* mini_keys; float* grad_send; int64_t n_positions;
} ScoreShardPlan;
int score_shard_plan(ScoreHandle h, int32_t world, ScoreShardPlan* out);
int score_shard_pack_grads(ScoreHandle h);
/* read-back of the all-gathered [world, world+1] count matrix (device pointer, n ints) into pinned memory of the handle:
 * fetch enqueues the copy + an event, wait blocks on that event only and copies the n ints to `out` - work enqueued in
 * between (the previous step's score_step_finish) keeps the device busy while the host learns the exchange sizes. */
int score_shard_counts_fetch(ScoreHandle h, const int32_t* counts_dev, int32_t n);
int score_shard_counts_wait(ScoreHandle h, int32_t* out, int32_t n);
int score_shard_presort(ScoreHandle h, const int32_t* ext_keys, int64_t n_ext);
/* Peer-memory exchange (NVLink / NVSwitch): gather and pack FUSED with the all-to-all.  count_matrix_dev = the all-gathered
 * [world][world+1] counts on the device; peers[r] = rank r's destination buffer as mapped into this process (device
 * addresses, e.g. torch symmetric memory).  score_shard_serve_push: LAZY catch-up of the served rows, then every served row is
 * stored straight into its requester's staged table (row slot+1, the layout of score_shard_plan).  score_shard_grad_push: every
 * gradient row is stored straight into its owner's gradient buffer, grouped by requester - the order score_step_finish
 * expects next to the served key list.  The caller runs a cross-rank barrier on score_stream() after each.
 * score_shard_register_staged: staged tables owned by the caller (two, alternating steps) whose addresses are stable, so the
 * half-step on them is replayed as a CUDA graph. */
int score_shard_serve_push(ScoreHandle h, const int32_t* want_dev, int64_t n_recv, const int32_t* count_matrix_dev,
                           int32_t world, int32_t rank, const uint64_t* peers);
int score_shard_grad_push(ScoreHandle h, const int32_t* count_matrix_dev, int32_t world, int32_t rank, const uint64_t* peers);
int score_shard_register_staged(ScoreHandle h, const float* a, const float* b);

/* Global index of the first sample of the batches this handle steps on (a data-parallel rank: rank * per-rank batch;
 * default 0).  It keys the dropout stream of tf.nn.dropout's stand-in (score.py:71-73), so N ranks draw the masks one
 * process would draw on the concatenated batch instead of N copies of the same mask. */
int score_set_sample_offset(ScoreHandle h, int32_t first_global_sample);

/* The CUDA stream every call of this handle is ordered on (a cudaStream_t). */
int score_stream(ScoreHandle h, void** cuda_stream);

/* Measurement hooks (bench.py).
 * score_launch_count: kernels launched by this library since the handle was created.
 * score_enable_probes / score_probe_times: CUDA-event timing of named kernels inside the timed
 *   train steps, on the stream they are launched on.  out[2*p] = accumulated ms, out[2*p+1] = number
 *   of samples, p = 0 coatt_fwd (gather), 1 coatt_bwd, 2 emb_update (scatter+Adam), 3 sort, 4 whole step,
 *   5 dense forward, 6 dense backward, 7 lazy catch-up.
 * score_last_step_stats: out3 = { positions N, live positions (non-zero key), unique rows U } of the
 *   last train step - the factors of the algorithmic byte counts in DESIGN.md. */
int64_t score_launch_count(ScoreHandle h);
int score_enable_probes(ScoreHandle h, int on);
int score_probe_times(ScoreHandle h, double* out, int n);
int score_last_step_stats(ScoreHandle h, int64_t* out3);

/* ------------------------------------------------------------------------------------------------------------
 * On-GPU graph store + neighbor sampler: replaces GraphHandler / GraphLoader
 * (code/score/graph_loader.py:94-277 neighbor selection, :340-385 batch assembly) for callers that keep the
 * interaction graph in HBM.  The graph is the content of the reference's per-node Mongo documents
 * (graph_storage.py:153-245) as two CSR arrays indexed by node * n_slices + slice; nodes are the unified ids
 * 1..n_user (users) and n_user+1..n_user+n_item (items), row 0 is unused.  All arrays are HOST pointers, copied once. */
typedef struct ScoreGraph* ScoreGraphHandle;
typedef struct ScoreGraphDesc {
    int32_t n_user, n_item, n_slices;
    int32_t user_fnum, item_fnum;  /* fields per node incl. the id (graph_loader.py:186-191)                   */
    const int64_t* hop1_off;       /* [(n_user+n_item+1)*n_slices + 1]  doc['1hop'][slice]                      */
    const int32_t* hop1_ids;
    const int64_t* hop2_off;       /* same shape                        doc['2hop'][slice] (<= 100 ids)         */
    const int32_t* hop2_ids;
    const int32_t* hop2_deg;       /* doc['degrees'][slice], parallel to hop2_ids; may be NULL ('rs' mode only) */
    const int32_t* user_feat;      /* [(n_user+1) * (user_fnum-1)]  user_feat_dict[str(uid)]; NULL if user_fnum == 1 */
    const int32_t* item_feat;      /* [(n_item+1) * (item_fnum-1)]  row iid - n_user; NULL if item_fnum == 1   */
} ScoreGraphDesc;
/* SCORE_ERR_ARG (before any device work) when an offset array is not ascending from 0 or a list holds an id of the wrong
   type: 1-hop neighbors of a user / 2-hop neighbors of an item are item ids (n_user+1 .. n_user+n_item), 1-hop neighbors
   of an item / 2-hop neighbors of a user are user ids (1 .. n_user), 0 is the dummy - graph_storage.py:127-246 builds the
   lists that way, and graph_loader.py:186-191 raises KeyError for an id without a feature entry */
int score_graph_create(const ScoreGraphDesc* desc, int device, ScoreGraphHandle* out);
int score_graph_destroy(ScoreGraphHandle g);
const char* score_graph_last_error(ScoreGraphHandle g);
/* One batch, as GraphLoader.worker assembles it (graph_loader.py:340-385): uids[ceil(B/group)] target users,
 * iids[B] target items (group = 1 + neg_sample_num consecutive items per user), histories of slices
 * start_time..pred_time-1 padded to max_time_len with copies of the last one.  mode 0 = 'rs' (uniform with
 * replacement, the mode train_score.py uses), 1 = 'is' (softmax of 1/(degree-1)).  uids / iids: host or device.
 * *out receives DEVICE pointers (on_device = 1) owned by the graph handle, valid until the next sample call; pass it
 * to score_train_step / score_eval.  Work is enqueued on cuda_stream (a cudaStream_t; NULL = the graph's own). */
int score_graph_sample(ScoreGraphHandle g, const int32_t* uids, const int32_t* iids, int32_t batch_size,
                       int32_t group, int32_t start_time, int32_t pred_time, int32_t max_time_len,
                       int32_t obj_per_time_slice, int32_t mode, uint64_t seed, uint32_t draw_id, void* cuda_stream,
                       ScoreBatch* out);
/* 2-hop graph construction: replaces GraphStore.construct_coll_2hop (code/graph_storage.py:127-246).
 * Input: the 1-hop lists of every node and slice as a CSR array (hop1_off / hop1_ids as in ScoreGraphDesc, HOST pointers).
 * Output (HOST pointers): hop1_ids_out - the same lists with the in-place shuffles of the lists longer than max_1hop applied
 * (graph_storage.py:171-173, 211-213: the new documents store the shuffled lists); hop2_off_out [(n_user+n_item+1)*n_slices+1],
 * hop2_ids_out / hop2_deg_out [*n2_out] - doc['2hop'] / doc['degrees'].  Call with hop2_ids_out == NULL to learn *n2_out
 * (hop2_off_out is filled), then again with buffers of that capacity.  Randomness: the k-th random.shuffle / np.random.choice
 * call of the reference's processing order (items, then users) becomes the stable argsort of Philox4x32-10 uniforms keyed
 * by (seed, stream 11 / 12, k, element) - csrc/hop2.cu; max_1hop <= 32. */
typedef struct ScoreHop2Desc {
    int32_t n_user, n_item, n_slices, start_time, max_1hop, max_2hop;
    const int64_t* hop1_off; const int32_t* hop1_ids;
    uint64_t seed;
} ScoreHop2Desc;
int score_graph_build_2hop(const ScoreHop2Desc* desc, int device, int32_t* hop1_ids_out, int64_t* hop2_off_out,
                           int32_t* hop2_ids_out, int32_t* hop2_deg_out, int64_t capacity, int64_t* n2_out);
const char* score_graph_build_2hop_error(void);
int score_graph_sync(ScoreGraphHandle g, void* cuda_stream);   /* waits; SCORE_ERR_ID_RANGE if a target was unknown */
int score_copy_to_host(void* dst_host, const void* src_device, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* SCORE_B200_H */
