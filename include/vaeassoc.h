/* libvaeassoc -- C-ABI of the B200-native associated-VAE train step.
 *
 * Drop-in boundary: the reference has no FFI layer; its model class funnels every
 * arithmetic op through ONE TensorFlow entry point, `self.sess.run(fetches, feed_dict)`
 * (/root/reference/vae_assoc.py:383,389,399,402,417,423).  Each function below replaces one of those
 * sess.run call sites (cited per function).  The Python class `vae_assoc_b200.vae_assoc.
 * AssocVariationalAutoEncoder` binds them with ctypes and keeps the reference's surface.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; `vaeassoc_last_error(h)` gives the text
 *     (pass NULL for errors of vaeassoc_create).  No C++ exception crosses the ABI.
 *   - a handle is bound to one CUDA device and one compute stream; calls on one handle are serialised by an
 *     internal mutex (the reference's callers drive the model from a Qt worker thread,
 *     baxter_vae_assoc_writer.py:654-674).
 *   - all tensors are fp32, row-major.  `*_dev` pointers are device pointers, `*_host` host pointers.
 *     The caller owns every I/O buffer; the library owns parameters, gradients, Adam slots, activations and
 *     workspaces.
 *   - there is NO CPU fallback: without a usable sm_100 device `vaeassoc_create` fails.
 */
#ifndef VAEASSOC_H_
#define VAEASSOC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAEASSOC_MAX_MODALITIES 4
#define VAEASSOC_ABI_VERSION 2

typedef struct vaeassoc_ctx* vaeassoc_handle;

enum { VAEASSOC_RELU = 0, VAEASSOC_SOFTPLUS = 1 };                 /* transfer_fct, vae_assoc.py:26,502   */
/* SIMT FFMA (1e-4 path) | tcgen05 kind::tf32 (2e-3 path) | bf16 operands: NAMED by SURVEY 8b's contract but REJECTED by
 * vaeassoc_create / vaeassoc_set_precision -- 8-bit mantissas miss the 2e-3 tolerance of the tensor-core path (tf32,
 * 10 bits, measures 6e-4), so the library refuses instead of silently training at another precision */
enum { VAEASSOC_FP32 = 0, VAEASSOC_TF32 = 1, VAEASSOC_BF16 = 2 };
enum { VAEASSOC_PARAMS = 0, VAEASSOC_GRADS = 1, VAEASSOC_ADAM_M = 2, VAEASSOC_ADAM_V = 3 };

/* one entry of `network_architectures` (vae_assoc.py:73-76) plus its `binary` / `weights` items (:31-41) */
typedef struct {
  int32_t n_input;
  int32_t n_hidden_recog_1, n_hidden_recog_2;
  int32_t n_hidden_gener_1, n_hidden_gener_2;   /* used by the conv variant only (vae_assoc.py:252,264,269,274) */
  int32_t hidden_conv;
  int32_t binary;
  float weight;
  char scope[32];               /* na["scope"] (vae_assoc.py:168,248: tf.variable_scope(scope)) -> variable names
                                   "<scope>/Variable_k", "<scope>_1/Variable_k"; empty -> "image","joint","modal2","modal3" */
} vaeassoc_modality;

/* constructor arguments of AssocVariationalAutoEncoder (vae_assoc.py:26-27) */
typedef struct {
  int32_t abi_version;          /* VAEASSOC_ABI_VERSION */
  int32_t n_modalities;
  int32_t batch_size;           /* rows THIS handle processes per step (static, vae_assoc.py:90) */
  int32_t n_z;
  int32_t transfer_fct;
  int32_t precision;
  int32_t device;               /* CUDA ordinal */
  int32_t use_graph;            /* replay the step as one CUDA graph */
  float assoc_lambda;
  float learning_rate;
  float beta1, beta2, adam_epsilon;   /* tf.train.AdamOptimizer defaults: 0.9, 0.999, 1e-8 */
  int64_t global_batch;         /* divisor of the batch-mean loss terms; 0 -> batch_size (single process) */
  int64_t global_row0;          /* first global sample index of this shard (Philox counters) */
  uint32_t eps_seed;
  uint32_t reserved;
  vaeassoc_modality mod[VAEASSOC_MAX_MODALITIES];
} vaeassoc_config;

/* logical tensor i (reference tf.Variable creation order) -> where it lives in the flat buffers */
typedef struct {
  char name[48];        /* TF-style variable name, e.g. "image/Variable_2", "image_1/Variable" */
  char role[16];        /* "W1","b1",...,"Vo","co" (oracle/vae_assoc_oracle.py header) */
  int32_t modality;
  int32_t ndim;         /* 1, 2 or 4 */
  int32_t shape[4];     /* logical shape, reference layout */
  int64_t offset;       /* float offset of element [0,0] in every flat buffer */
  int64_t rows, cols;   /* 2-D view: rows x cols (4-D conv filters: rows = kh*kw*d2, cols = d3) */
  int64_t ld;           /* physical row stride (floats) */
} vaeassoc_tensor_info;

/* ---- lifetime ------------------------------------------------------------------------------------------ */
/* replaces AssocVariationalAutoEncoder.__init__'s graph build + tf.InteractiveSession (vae_assoc.py:54-67) */
int vaeassoc_create(const vaeassoc_config* cfg, vaeassoc_handle* out);
int vaeassoc_destroy(vaeassoc_handle h);
const char* vaeassoc_last_error(vaeassoc_handle h);
int vaeassoc_abi_version(void);
/* run the handle's work on an existing stream (e.g. torch's current stream); NULL -> the handle's own */
int vaeassoc_set_stream(vaeassoc_handle h, void* cuda_stream);
int vaeassoc_stream_sync(vaeassoc_handle h);
int vaeassoc_set_precision(vaeassoc_handle h, int precision);
int vaeassoc_set_learning_rate(vaeassoc_handle h, float lr);

/* ---- parameters / optimiser state (replaces tf.Variable access; tf.train.Saver, vae_assoc.py:70) --------- */
int vaeassoc_num_tensors(vaeassoc_handle h);
int vaeassoc_layout_query(vaeassoc_handle h, int i, vaeassoc_tensor_info* out);
int64_t vaeassoc_flat_size(vaeassoc_handle h);                       /* floats per flat buffer (padded)  */
void* vaeassoc_flat_ptr(vaeassoc_handle h, int which);               /* device pointer of a flat buffer  */
int vaeassoc_tensor_set(vaeassoc_handle h, int which, int i, const float* src_host);   /* dense, logical */
int vaeassoc_tensor_get(vaeassoc_handle h, int which, int i, float* dst_host);
int vaeassoc_step_get(vaeassoc_handle h, int64_t* step);             /* Adam step count t                */
int vaeassoc_step_set(vaeassoc_handle h, int64_t step);

/* ---- the hot path ------------------------------------------------------------------------------------------
 * vaeassoc_train_step  replaces  sess.run((self.optimizer, self.cost), feed_dict)   vae_assoc.py:383-384
 *   x_dev[m]  : [batch_size, n_input_m] device rows with leading dimension ld[m] (floats; NULL -> n_input_m)
 *   eps_dev   : [batch_size, n_z] injected reparameterisation noise, or NULL -> Philox4x32-10
 *               (key (eps_seed, 1), counter (global row, block, step)) replacing tf.random_normal :90
 * The step is asynchronous; the scalar cost of step t lands in a device-side history ring and is fetched with
 * vaeassoc_cost_read (which synchronises), so a training loop need not sync every step as the reference does.
 * vaeassoc_grad_step = the same without the Adam update (gradients stay readable through VAEASSOC_GRADS).  After a
 * vaeassoc_train_step the gradient buffer is unspecified: on one GPU the Adam kernel hands the accumulator back cleared
 * (the next step then needs no memset), like the gradient tensors of the reference's `minimize` op, which no caller
 * can fetch either (vae_assoc.py:373-374). */
int vaeassoc_train_step(vaeassoc_handle h, const float* const* x_dev, const int64_t* ld, const float* eps_dev);
int vaeassoc_grad_step(vaeassoc_handle h, const float* const* x_dev, const int64_t* ld, const float* eps_dev);
int vaeassoc_adam_step(vaeassoc_handle h);                            /* ApplyAdam on the current gradients */
int vaeassoc_cost_read(vaeassoc_handle h, float* cost_host);          /* cost of the most recent step       */
int vaeassoc_cost_history(vaeassoc_handle h, int64_t first_step, int64_t n, float* dst_host);
/* host-buffer form of partial_fit (vae_assoc.py:378-386): H2D copy of X, one step, D2H of the cost.
 * x_host[m] dense [batch_size, n_input_m]; eps_host may be NULL. */
int vaeassoc_partial_fit_host(vaeassoc_handle h, const float* const* x_host, const float* eps_host,
                              float* cost_host);
/* pipelined host form used by train(): uploads batch k+1 on a copy stream while step k computes.
 * `vaeassoc_submit_host` returns as soon as the batch is queued; costs are read back with
 * vaeassoc_cost_history / vaeassoc_cost_read.
 * BUFFER LIFETIME: the H2D copies are asynchronous.  Pageable buffers are staged by the driver before the call returns
 * and may be reused at once; PINNED (page-locked) buffers are read by DMA after the call returns and must stay
 * allocated and unmodified until `vaeassoc_upload_wait(h, k)` has returned for this submit's index k (= the number of
 * submits made before it; vaeassoc_submit_count).  train() and bench.py rotate pinned slots behind that call. */
int vaeassoc_submit_host(vaeassoc_handle h, const float* const* x_host, const float* eps_host);
/* device-resident data set (replaces dataset.py:22-43 next_batch + the column slicing of vae_assoc.py:510,543 + the
 * feed_dict copy): `data_dev` = the whole training matrix [n_rows, >= sum n_input] uploaded ONCE (row pitch `ld` floats;
 * modality m occupies columns [sum_{k<m} n_input_k, +n_input_m), the reference's sens_indices); `index_host` = the
 * batch_size row indices of this step (the epoch's shuffled order).  Only the indices cross PCIe (8 B per pair). */
int vaeassoc_submit_indexed(vaeassoc_handle h, const float* data_dev, int64_t ld, int64_t n_rows,
                            const int64_t* index_host, const float* eps_host);
int64_t vaeassoc_submit_count(vaeassoc_handle h);                     /* submits so far = index of the next one */
int vaeassoc_upload_wait(vaeassoc_handle h, int64_t submit_index);    /* blocks until that submit's H2D completed */
/* every submit also queues an async D2H of that step's cost into a pinned host ring (4096 entries);
 * this synchronises the stream and returns the costs of submits [first_submit, first_submit + n). */
int vaeassoc_submit_costs(vaeassoc_handle h, int64_t first_submit, int64_t n, float* dst_host);

/* evaluate_cost (vae_assoc.py:388-391): forward + loss, no gradients, no update */
int vaeassoc_eval_cost(vaeassoc_handle h, const float* const* x_dev, const int64_t* ld, const float* eps_dev,
                       float* cost_host);
/* transform (vae_assoc.py:393-403): z_mean (and log sigma^2) of one modality; outputs dense [B, n_z] device */
int vaeassoc_encode(vaeassoc_handle h, int modality, const float* x_dev, int64_t ld, float* mu_dev,
                    float* logvar_dev);
/* generate (vae_assoc.py:405-419): feeds z directly; xhat_dev dense [B, n_input_m] device */
int vaeassoc_decode(vaeassoc_handle h, int modality, const float* z_dev, float* xhat_dev);
/* reconstruct (vae_assoc.py:421-425): encode + sample + decode of ONE modality */
int vaeassoc_reconstruct(vaeassoc_handle h, int modality, const float* x_dev, int64_t ld, const float* eps_dev,
                         float* xhat_dev);
/* The same three entry points for HOST callers (what baxter_vae_assoc_writer.py:142-145 and vae_assoc_model_viewer.py:113
 * pass: numpy arrays): ONE call = the inputs' H2D, the forward launches of every requested modality, packing and ONE
 * D2H, captured as one CUDA graph per (kind, modality set); blocking.  kind 0 transform: x_host[m] -> out_host[m] =
 * z_mean [B, n_z]; kind 1 generate: z_or_eps_host = z [B, n_z] -> out_host[m] = x_reconstr_mean [B, n_input_m];
 * kind 2 reconstruct: x_host[m] (+ optional injected eps [B, n_z], NULL = Philox) -> out_host[m].  modality -1 = all
 * modalities (dense tf32 models then run the fused encoder / decoder / whole-forward launch of the tile kernel). */
int vaeassoc_infer_host(vaeassoc_handle h, int kind, int modality, const float* const* x_host, const float* z_or_eps_host,
                        float* const* out_host);

/* the probe points of vae_assoc.py:545-571 after the most recent step.  kind: */
enum {
  VAEASSOC_PROBE_Z_MEAN = 0,        /* [B, n_z]  per modality          vae_assoc.py:114 */
  VAEASSOC_PROBE_Z_LOG_SIGMA_SQ,    /* [B, n_z]                        :115 */
  VAEASSOC_PROBE_Z,                 /* [B, n_z]                        :116 */
  VAEASSOC_PROBE_X_RECONSTR_MEAN,   /* [B, n_input_m]                  :117 */
  VAEASSOC_PROBE_RECONSTR_LOSS,     /* [B] (binary) or [1] (Gaussian)  :338 */
  VAEASSOC_PROBE_LATENT_LOSS,       /* [B]                             :339 */
  VAEASSOC_PROBE_VAE_COST,          /* [1]                             :340 */
  VAEASSOC_PROBE_ASSOC_COST,        /* [1]  sum over modality pairs    :344-366 (modality ignored) */
  VAEASSOC_PROBE_D_Z_MEAN,          /* [B, n_z]  d cost / d z_mean     (autodiff of :373-374) */
  VAEASSOC_PROBE_D_Z_LOG_SIGMA_SQ,  /* [B, n_z] */
  VAEASSOC_PROBE_EPS                /* [B, n_z]  the noise the step used (modality ignored) */
};
int vaeassoc_probe_get(vaeassoc_handle h, int kind, int modality, float* dst_host, int64_t capacity_floats,
                       int64_t* n_written);
/* relu sign bits of a hidden layer after the most recent tensor-core step (tf32 mode, relu, dense modality): layer
 * 0 = h1, 1 = h2 (encoder), 2 = g1, 3 = g2 (decoder); bit (row, col) = activation > 0, 32 columns per word,
 * `*words_per_row` words per row.  These are the masks the dgrad epilogues apply (TF ReluGrad, y > 0); parity tests feed
 * them to the oracle so that gradients are compared under IDENTICAL masks.  Fails if the layer has no mask. */
int vaeassoc_probe_mask(vaeassoc_handle h, int layer, int modality, uint32_t* dst_host, int64_t capacity_words,
                        int64_t* n_written, int64_t* words_per_row);

/* ---- synthetic paired batches: replaces dataset.py:22-43 next_batch + utils.py:142-195 ------------------- */
/* rows [row0, row0 + n_rows) of the global stream; x_dev[m] dense [n_rows, n_input_m] device */
int vaeassoc_synth_batch(vaeassoc_handle h, uint32_t data_seed, uint32_t proj_seed, int64_t row0, int64_t n_rows,
                         float* const* x_dev);
/* Philox N(0,1) rows, exposed for tests and for generate(z_mu=None) (vae_assoc.py:414) */
int vaeassoc_philox_normal(vaeassoc_handle h, uint32_t seed, uint32_t tag, int64_t row0, int64_t n_rows,
                           int32_t n_cols, uint32_t step, float* dst_dev);

/* ---- data parallelism (new; the reference is single process) ------------------------------------------------
 * One NCCL all-reduce(SUM) per gradient bucket on a communication stream, overlapped with the rest of the
 * backward pass; the scalar cost rides in a spare slot of the flat gradient buffer. */
int vaeassoc_comm_unique_id(const char* nccl_lib_path, void* id128);           /* rank 0: ncclGetUniqueId   */
int vaeassoc_comm_init(vaeassoc_handle h, const char* nccl_lib_path, const void* id128, int rank, int world);
int vaeassoc_comm_destroy(vaeassoc_handle h);
/* broadcast rank 0's parameters, Adam slots and step count to every rank (replicas must start identical: the
 * all-reduce only averages gradients); collective -- every rank calls it */
int vaeassoc_comm_sync_state(vaeassoc_handle h);
/* ncclCommGetAsyncError: 0 = healthy; on an asynchronous NCCL failure the communicator is aborted (ncclCommAbort),
 * the handle's error text is set and a non-zero status is returned (also checked by vaeassoc_stream_sync) */
int vaeassoc_comm_check(vaeassoc_handle h);

/* ---- peer-memory data-parallel step (one NVSwitch box, <= 8 ranks) ---------------------------------------------------
 * Replaces `ncclAllReduce` + the replicated Adam of the schedule above by ONE kernel per step (csrc/peer_adam.cu):
 * reduce-scatter of the flat gradient buffers read straight from the peers' HBM over NVLink, Adam on the owned shard,
 * all-gather of the updated parameters (+ tf32 shadow) by peer stores; cross-GPU arrival words instead of a host or
 * NCCL barrier.  Set-up: every rank exports one blob (cudaIpc handle of the block holding its flat buffers), the host
 * side all-gathers the blobs (torch.distributed, vae_assoc.py: init_data_parallel) and every rank attaches ALL of them
 * (rank order), then the ranks barrier once before the first step.  Needs vaeassoc_comm_init first (rank / world; NCCL
 * still serves vaeassoc_comm_sync_state, evaluate_cost and compute_gradients).  Adam's m / v become sharded: reads
 * through vaeassoc_tensor_get / vaeassoc_save pull the peers' shards first.  vaeassoc_peer_detach (or comm_destroy)
 * must be preceded by a barrier of the ranks: a peer may not touch a detached rank's memory. */
#define VAEASSOC_PEER_BLOB_BYTES 128
int vaeassoc_peer_export(vaeassoc_handle h, void* blob /* VAEASSOC_PEER_BLOB_BYTES */);
int vaeassoc_peer_attach(vaeassoc_handle h, const void* all_blobs /* world x VAEASSOC_PEER_BLOB_BYTES, rank order */);
/* The same step over SYMMETRIC memory (e.g. torch.distributed._symmetric_memory: identical allocation on every rank,
 * mapped into every process, optionally bound to an NVSwitch multicast object): the library moves its flat buffers
 * (vaeassoc_arena_floats floats) into `local_arena`, reaches rank r's through rank_arenas[r] and, when `multicast_arena`
 * is non-null, reduces the gradients INSIDE the switch (multimem.ld_reduce) and broadcasts the updated parameters with one
 * store per element (multimem.st) -- 1/world of the NVLink bytes of the plain peer form.  The caller keeps the memory
 * alive until vaeassoc_destroy. */
int64_t vaeassoc_arena_floats(vaeassoc_handle h);
int vaeassoc_peer_attach_symmetric(vaeassoc_handle h, void* local_arena, const void* const* rank_arenas, void* multicast_arena);
int vaeassoc_peer_multicast(vaeassoc_handle h);   /* 1 while the NVLS (multimem) form is in use */
int vaeassoc_peer_detach(vaeassoc_handle h);
int vaeassoc_peer_active(vaeassoc_handle h);      /* 1 while the peer-memory step is in use */

/* Own bounds check (debug): every device buffer of the handle ends in a 256-byte guard pattern that no kernel may
 * write; reports how many guards exist and how many no longer hold the pattern (0 expected, tests/test_gpu_parity.py). */
int vaeassoc_debug_guard_check(vaeassoc_handle h, int64_t* n_guards, int64_t* n_corrupt);

/* ---- checkpoints: tf.train.Saver.save / .restore over ALL variables incl. the Adam slots (vae_assoc.py:70,427-463) --
 * One self-describing binary file ("VAEASSOC" magic, tensor table with the TF-style names, parameters, both Adam
 * slots, step count).  vaeassoc_load matches tensors by NAME and shape and fails (non-zero, handle unchanged) on a
 * mismatch.  TensorFlow V1 checkpoints are imported on the Python side (vae_assoc_b200/tf_checkpoint.py). */
int vaeassoc_save(vaeassoc_handle h, const char* path);
int vaeassoc_load(vaeassoc_handle h, const char* path);

/* ---- introspection for benchmarks --------------------------------------------------------------------------- */
/* number of kernels this library launched on the handle since creation (bench.py's gpu_launches) */
int64_t vaeassoc_launch_count(vaeassoc_handle h);
/* per-op timing of one eager (non-graph) step with CUDA events: names[i] (<=32 chars) and ms[i]; returns the
 * number of ops written (<= capacity) or a negative error */
int vaeassoc_profile_step(vaeassoc_handle h, const float* const* x_dev, const int64_t* ld, const float* eps_dev,
                          char* names, float* ms, double* flops, double* bytes, int capacity);

/* test hook: run ONE dense-layer contraction of the library on caller buffers (device pointers) and synchronise.
 * kind 0: C = act(A[M,K] B[K,N] + bias)   1: C = (A[M,K] B[N,K]^T) * act'(aux)   2: C += A[K,M]^T B[K,N],
 * bias_grad += colsum(B).  use_tc = 1: the tcgen05 kernel (fails if the shape is not served); 0: the fp32 kernel the
 * library would pick (skinny HBM-bound kernel when one extent <= 16, else the tiled SIMT kernel); 2: tiled SIMT. */
int vaeassoc_debug_gemm(vaeassoc_handle h, int kind, int use_tc, int M, int N, int K, const float* A, int64_t lda,
                        const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias, float* bias_grad,
                        const float* aux, int64_t ldaux, int act, int round_out);

#ifdef __cplusplus
}
#endif
#endif /* VAEASSOC_H_ */
