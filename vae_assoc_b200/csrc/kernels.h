// Internal launcher declarations of libvaeassoc (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vaeassoc {

// ------------------------------------------------------------------------------------------------------
// dense layers.  Three contractions cover forward / dgrad / wgrad of y = act(x W + b)  (vae_assoc.py:187-221,
// 259-303 and their autodiff, :373-374):
//   NN  C[M,N]  = act(A[M,K] . B[K,N] + bias[N])                     forward
//   NT  C[M,N]  = (A[M,K] . B[N,K]^T) (*) act'(aux[M,N])             dgrad (aux = stored activation of the layer below)
//   TN  C[M,N] += A[K,M]^T . B[K,N];  bias_grad[N] += colsum(B)      wgrad (K = batch; split-K: partial tiles in a
//                                                                     workspace, summed in a fixed order -> deterministic)
// ------------------------------------------------------------------------------------------------------
struct GemmArgs {
  int M = 0, N = 0, K = 0;
  const float* A = nullptr; int64_t lda = 0;
  const float* B = nullptr; int64_t ldb = 0;
  float* C = nullptr;       int64_t ldc = 0;
  const float* bias = nullptr;
  float* bias_grad = nullptr;
  const float* aux = nullptr; int64_t ldaux = 0;
  int act = 0;          // Act code: forward activation, or which act' to apply in dgrad
  int round_out = 0;    // round stored outputs to tf32 (they feed a tcgen05 kind::tf32 GEMM)
  int splitk = 1;       // TN only
  float* ws = nullptr;  // TN, SIMT / skinny kernels: workspace of gemm_*_ws_floats() floats for the split partial sums
  // relu layers on the tcgen05 path: the forward epilogue also writes one bit per output (activation > 0), 32 columns
  // per word, row pitch ldmask words; the dgrad epilogue reads these words instead of the fp32 activation tile (`aux`)
  uint32_t* mask_out = nullptr;
  const uint32_t* mask_in = nullptr;
  int64_t ldmask = 0;
  // NN on the tcgen05 path, output layer of a decoder: the epilogue turns the pre-activation straight into the
  // reconstruction loss and d cost / d pre-activation (vae_assoc.py:321-328): C receives d a, `loss_x` is the target,
  // every epilogue warp leaves its loss sum in loss_partials[((m_blk * tiles_n + n_blk) * 2 + CTA) * warps + warp]
  const float* loss_x = nullptr; int64_t ld_loss_x = 0;
  float* loss_partials = nullptr;
  float loss_scale = 0.f;   // binary: w / B_global ; Gaussian: w
  int loss_binary = 0;
  // NN / NT on the tcgen05 path: the tile's raw accumulator is ADDED into C (TMA reduce-add; no bias, activation,
  // rounding) -- a contraction with a tiny N and a long K is cut into k ranges that run as parallel tasks
  int force_reduce = 0;
};

void launch_gemm_nn_simt(const GemmArgs& a, cudaStream_t s);
void launch_gemm_nt_simt(const GemmArgs& a, cudaStream_t s);
void launch_gemm_tn_simt(const GemmArgs& a, cudaStream_t s);
int64_t gemm_tn_simt_ws_floats(const GemmArgs& a);
// out[cols] += column sums of X[rows, cols] (bias gradient when the wgrad GEMM runs on the tensor cores); `ws` =
// colsum_ws_floats(rows, cols) zero-initialised floats owned by the call site (per-CTA partials + arrival tickets; the
// kernel leaves the tickets at zero), so that the sum has a fixed order
int64_t colsum_ws_floats(int64_t rows, int cols);
void launch_colsum(const float* X, int64_t ld, int64_t rows, int cols, float* out, float* ws, cudaStream_t s);

// HBM-bound kernels for contractions with one extent <= 16 (the n_z-wide heads / decoder input layer), gemm_skinny.cu
bool skinny_supported(int kind /*0 NN,1 NT,2 TN*/, const GemmArgs& a);
void launch_gemm_skinny(int kind, const GemmArgs& a, cudaStream_t s);
int64_t gemm_skinny_ws_floats(int kind, const GemmArgs& a);

// tcgen05 / TMA path (gemm_group.cu): a persistent kernel that executes a list of 256 x BN tile tasks drawn from
// several contractions ("problems"), ordered by dependency and linked by row-block completion counters.  A plan owns
// the problems / tasks of one handle; a launch runs a contiguous task range.  Shapes the path cannot serve fall back
// to the SIMT *kernels of this library*, never to a CPU or a vendor library.
struct GroupPlan;
bool tc_supported(int kind /*0 NN,1 NT,2 TN*/, const GemmArgs& a);
GroupPlan* group_create();
void group_destroy(GroupPlan* g);
int group_add_problem(GroupPlan* g, int kind, const GemmArgs& a, char* err, int errlen);   // problem index or -1
int group_problem_tiles_m(const GroupPlan* g, int prob);
int group_problem_tiles_n(const GroupPlan* g, int prob);
int group_problem_kblocks(const GroupPlan* g, int prob);      // k-blocks (32 deep; 64 at batches of one to three row blocks)
int group_problem_kb_per_rowblock(const GroupPlan* g, int prob);   // k-blocks per 256 batch rows of a weight gradient (8 or 4)
// operands ready when counters[wait_ctr .. +wait_cnt) >= wait_val (and counters[wait2_ctr] >= wait2_val); every
// epilogue warp (16 per tile) bumps counters[signal_ctr] once the tile is globally visible; -1 / 0 = none
// extra_flags: kTaskHalf (consumer of half-tile hand-overs: wait2_ctr = half counter, wait2_val = k-blocks per producing
// tile, the half counter's target is wait_val as well) | kTaskSigHalf (producer: publishes its first half early)
constexpr int kTaskHalf = 4096, kTaskSigHalf = 8192;
int group_add_task(GroupPlan* g, int prob, int m_blk, int n_blk, int kb0, int nkb, int wait_ctr, int wait_cnt,
                   int wait_val, int wait2_ctr, int wait2_val, int signal_ctr, int extra_flags = 0);
int group_problem_bn(const GroupPlan* g, int prob);
void group_set_deep_k(GroupPlan* g, bool allow);   // call before adding problems
int group_num_tasks(const GroupPlan* g);
// elementwise task over the 256 rows of row block m_blk, executed by the epilogue warps of whichever CTA pair pops it
// (kind 0: latent forward, 1: latent backward -- arguments from group_set_elem); waits / signals like a tile task
// kind 2: cost finalize (sums the block partials of the loss / latent tasks; arguments: GElem::fin), one per launch
int group_add_elt_task(GroupPlan* g, int kind, int m_blk, int batch, int wait_ctr, int wait_cnt, int wait_val,
                       int wait2_ctr, int wait2_val, int signal_ctr, int wait2_cnt = 1, int variant = 0);
struct GElem;
void group_set_elem(GroupPlan* g, const GElem& e);
// a launch site = the problems / tasks added between group_begin() and group_end(): ONE kernel launch (<= 24 problems)
int group_begin(GroupPlan* g);
bool group_end(GroupPlan* g, char* err, int errlen);
int group_site_tasks(const GroupPlan* g, int site);
// half_off: the half-tile counter of counter i is counter i + half_off (0: the plan uses none)
void group_set_counters(GroupPlan* g, uint32_t* d_counters, int n, int half_off = 0);
bool group_upload(GroupPlan* g, char* err, int errlen);
// dynamic_first: every task (the first one included) comes from the atomic queue -- required whenever a kernel that
// waits on a peer GPU (NCCL) may hold SMs while this launch is resident
// advance: the finalize task of the launch (if any) bumps the Adam step counter
void group_launch(const GroupPlan* g, int site, uint32_t* queue, int reset_first, int reset_count, int dynamic_first,
                  cudaStream_t s, int advance = 0);
void group_debug_timeline(const GroupPlan* g, int site, uint32_t* queue, int reset_first, int reset_count,
                          cudaStream_t s);   // debug: per-task %globaltimer stamps to stderr
#ifndef VAEASSOC_EPI_WARPS
#define VAEASSOC_EPI_WARPS 8            // epilogue warps per CTA of the tile kernel (build-time variant: 16)
#endif
constexpr int kGroupSignalsPerTile = 2 * VAEASSOC_EPI_WARPS;   // epilogue warps of a CTA pair
constexpr bool kGroupHalfOk = VAEASSOC_EPI_WARPS == 8;         // half-tile hand-overs need two epilogue slots per lane quarter

// ------------------------------------------------------------------------------------------------------
// conv / transposed-conv layers of the hidden_conv=True modality as im2col -> GEMM -> col2im (conv.cu)
// ------------------------------------------------------------------------------------------------------
struct Im2colArgs {
  const float* x = nullptr;          // [B, H, W, C] dense NHWC
  int B = 0, H = 0, W = 0, C = 0;
  int k = 0, s = 1, pb = 0;          // kernel, stride, pad_before (TensorFlow SAME/VALID rule)
  int OH = 0, OW = 0;                // conv output size = number of patch rows per image
  float* out = nullptr; int64_t ldo = 0;   // [B*OH*OW, k*k*C] with row pitch ldo
};
void launch_im2col(const Im2colArgs& a, cudaStream_t s);

struct Col2imArgs {
  const float* cols = nullptr; int64_t ldc = 0;   // [B*h*w, k*k*C]
  int B = 0, H = 0, W = 0, C = 0;    // output [B, H, W, C]
  int h = 0, w = 0;                  // spatial size of the column rows
  int k = 0, s = 1, pb = 0;
  const float* bias = nullptr;       // [C] or null
  int act = 0, round_out = 0;
  float* out = nullptr;
};
void launch_col2im(const Col2imArgs& a, cudaStream_t s);

// ------------------------------------------------------------------------------------------------------
// input staging + noise
// ------------------------------------------------------------------------------------------------------
struct StageArgs {
  int n_mod = 0;
  int batch = 0;
  const float* src[4] = {nullptr, nullptr, nullptr, nullptr};
  int64_t src_ld[4] = {0, 0, 0, 0};
  const int64_t* row_index = nullptr;  // device-resident data set: batch row r is source row row_index[r] (null: r)
  float* dst[4] = {nullptr, nullptr, nullptr, nullptr};
  int64_t dst_ld[4] = {0, 0, 0, 0};
  int n_input[4] = {0, 0, 0, 0};
  int round_tf32 = 0;
  // eps: copy `eps_src` (dense [B, n_z]) or generate with Philox when null
  const float* eps_src = nullptr;
  float* eps_dst = nullptr;
  int n_z = 0;
  uint32_t eps_seed = 0;
  int64_t global_row0 = 0;
  const int64_t* step_dev = nullptr;   // Adam step counter t (eps of the step about to run uses t)
  // accumulators of the step's split-K contractions (heads, d z), cleared here so that no extra launch is needed
  float* zero_ptr[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int64_t zero_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // floats (multiples of 4)
};
void launch_stage(const StageArgs& a, cudaStream_t s);
void launch_philox_normal(float* dst, int64_t n_rows, int n_cols, uint32_t seed, uint32_t tag, int64_t row0,
                          uint32_t step, cudaStream_t s);

// ------------------------------------------------------------------------------------------------------
// fused reparameterisation + losses (vae_assoc.py:102-103, 319-371)
// ------------------------------------------------------------------------------------------------------
constexpr int kMaxPartialBlocks = 1184;   // 8 x 148
constexpr int kCostSlots = 16;            // per-block partial sums: [2m] recon, [2m+1] latent, [8] assoc

struct LatentArgs {
  int n_mod = 0, batch = 0, n_z = 0;
  float inv_global_batch = 0.f;     // 1 / B_global  (mean terms)
  float lambda = 0.f;               // assoc_lambda
  float weight[4] = {0, 0, 0, 0};
  const float* heads[4] = {nullptr, nullptr, nullptr, nullptr};   // [B, 2 n_z] = (mu | log sigma^2), ld = 2 n_z
  const float* eps = nullptr;                                     // [B, n_z]
  float* z[4] = {nullptr, nullptr, nullptr, nullptr};             // [B, n_z]
  float* gstat[4] = {nullptr, nullptr, nullptr, nullptr};         // [B, 2 n_z] prior-KL + assoc-KL gradient part
  float* latent_loss[4] = {nullptr, nullptr, nullptr, nullptr};   // [B]
  float* partials = nullptr;                                      // [kMaxPartialBlocks][kCostSlots]
  int with_grad = 1;
  int round_z = 0;                  // z feeds a tcgen05 GEMM (n_z large enough): round to tf32
  // split-K heads layer (tile-kernel tasks only): `heads` holds the bias-free sums of the k ranges; the latent task adds
  // the layer's bias (mu | log sigma^2, 2 n_z values) and writes the completed row back
  const float* head_bias[4] = {nullptr, nullptr, nullptr, nullptr};
};
int launch_latent_fwd(const LatentArgs& a, cudaStream_t s);       // returns number of blocks (partials rows)

struct LatentBwdArgs {
  int n_mod = 0, batch = 0, n_z = 0;
  const float* heads[4] = {nullptr, nullptr, nullptr, nullptr};
  const float* gstat[4] = {nullptr, nullptr, nullptr, nullptr};
  const float* dz[4] = {nullptr, nullptr, nullptr, nullptr};      // [B, n_z] gradient arriving from the decoder
  const float* eps = nullptr;
  float* dheads[4] = {nullptr, nullptr, nullptr, nullptr};        // [B, 2 n_z] = (d mu | d log sigma^2)
  int round_out = 0;
};
void launch_latent_bwd(const LatentBwdArgs& a, cudaStream_t s);

struct FinalizeArgs {
  int n_mod = 0;
  int binary[4] = {0, 0, 0, 0};
  float weight[4] = {0, 0, 0, 0};
  float inv_global_batch = 0.f, lambda = 0.f;
  const float* partials_latent = nullptr; int blocks_latent = 0;
  // reconstruction partials of modality m: blocks_recon[m] values at partials_recon[m][i * stride_recon[m] + off_recon[m]]
  const float* partials_recon[4] = {nullptr, nullptr, nullptr, nullptr}; int blocks_recon[4] = {0, 0, 0, 0};
  int stride_recon[4] = {kCostSlots, kCostSlots, kCostSlots, kCostSlots}; int off_recon[4] = {0, 2, 4, 6};
  float* scalars = nullptr;         // [16]: [m] vae_cost_m, [4+m] recon sum, [8] assoc sum, [9] cost (local)
  float* cost_slot = nullptr;       // spare slot of the flat gradient buffer (all-reduced with the grads)
  int64_t* step_dev = nullptr;      // incremented when `advance`
  int advance = 0;
};

// arguments of the elementwise tasks of a tile-kernel plan (kernel parameter of gemm_group_kernel)
struct GElem {
  LatentArgs lf;             // partials: [(row block * 2 + CTA) * 4 + warp][kCostSlots]
  LatentBwdArgs lb;
  float* bh_grad[4] = {nullptr, nullptr, nullptr, nullptr};   // bias gradient of the heads layer per modality
                                                              // (column sums of d mu | d log sigma^2), or null
  FinalizeArgs fin;          // one-launch form: the cost reduction is a task of the tile kernel as well
};

struct ReconArgs {
  int batch = 0, n_input = 0, binary = 0, slot = 0;
  float scale = 0.f;                 // binary: w / B_global ; Gaussian: w
  const float* x = nullptr; int64_t ldx = 0;
  const float* xhat = nullptr; int64_t ldxh = 0;   // sigmoid already applied by the GEMM epilogue when binary
  float* da = nullptr; int64_t ldda = 0;           // d cost / d pre-activation (null: loss only)
  float* row_loss = nullptr;                       // [B] (binary) ; null otherwise
  float* partials = nullptr;                       // this kernel's rows of [blocks][kCostSlots]
  int round_tf32 = 0;
};
int launch_recon_loss(const ReconArgs& a, cudaStream_t s);        // returns number of blocks

void launch_finalize(const FinalizeArgs& a, cudaStream_t s);

// ------------------------------------------------------------------------------------------------------
// Adam (TensorFlow ApplyAdam formulation, vae_assoc.py:373-374) over the flat buffers
// ------------------------------------------------------------------------------------------------------
struct AdamArgs {
  float* p = nullptr; const float* g = nullptr; float* m = nullptr; float* v = nullptr;
  float* p_tf32 = nullptr;          // rounded shadow copy read by the tcgen05 GEMMs (null in fp32 mode)
  int64_t n = 0;                    // floats (multiple of 4)
  float lr = 0.f, beta1 = 0.f, beta2 = 0.f, eps = 0.f;
  const int64_t* step_dev = nullptr;
  const float* cost_slot = nullptr; float* cost_hist = nullptr; int hist_cap = 0;   // publish cost of step t
  float* last_cost = nullptr;
  float* zero_g = nullptr;          // = g: clear the gradients once consumed (single-GPU one-launch schedule), or null
};
void launch_adam(const AdamArgs& a, cudaStream_t s);

// Data-parallel step over NVLink peer memory (peer_adam.cu): reduce-scatter of the flat gradient buffers + Adam on the
// owned shard + all-gather of the updated parameters, ONE kernel, no NCCL on the step.  Rank r owns float4 indices
// [shard_lo, shard_hi) of the flat buffers; every pointer table is indexed by rank (entry `rank` = the local buffer).
constexpr int kMaxPeers = 8;
struct PeerAdamArgs {
  AdamArgs adam;                     // local p / m / v / p_tf32, hyper-parameters, cost publication
  int world = 1, rank = 0;
  const float* g_peer[kMaxPeers] = {};   // flat gradient buffer of every rank (n + 32 floats; [n] = cost slot)
  float* p_peer[kMaxPeers] = {};
  float* ptf_peer[kMaxPeers] = {};       // tf32 shadow of every rank (entries null in fp32 mode)
  uint32_t* flag_peer[kMaxPeers] = {};   // [2][kMaxPeers] arrival words of every rank: [phase][source rank] = epoch
  // NVLS (NVSwitch multicast) form: one address that names the same offset of EVERY rank's arena.  multimem.ld_reduce on it
  // returns the sum over all ranks computed inside the switch (1/world of the inbound bytes); multimem.st on it stores to
  // every rank at once (1/world of the outbound bytes).  Null: plain peer loads / stores through the tables above.
  const float* g_mc = nullptr;
  float* p_mc = nullptr;
  float* ptf_mc = nullptr;
  uint32_t* sync = nullptr;              // local: [0] epoch of the last completed step, [1] CTAs done
  int64_t shard_lo = 0, shard_hi = 0;    // float4 units
  unsigned long long* tl = nullptr;      // debug (VAEASSOC_PEER_TIMELINE): [0] steps, [1..4] summed ns of the phases
};
void launch_peer_adam(const PeerAdamArgs& a, cudaStream_t s);
// waits until every rank's phase-B arrival word ([1][r] of `flags`) has reached the epoch of the last completed step
// (sync[0]): the peers' parameter stores have landed in this rank's memory and the peers have read its gradients
void launch_peer_wait(const uint32_t* flags, const uint32_t* sync, int world, cudaStream_t s);
void launch_round_copy(const float* src, float* dst, int64_t n, cudaStream_t s);   // dst = round_tf32(src)
void launch_publish_cost(const float* cost_slot, float* last_cost, cudaStream_t s);

// ------------------------------------------------------------------------------------------------------
// synthetic paired batches (replaces dataset.py / utils.py)
// ------------------------------------------------------------------------------------------------------
void launch_synth_projection(float* P, float* inv_std, int modality, int n_input, uint32_t proj_seed,
                             cudaStream_t s);
void launch_synth_modality(float* x, int64_t ldx, const float* P, const float* inv_std, int modality, int n_input,
                           int binary, uint32_t data_seed, int64_t row0, int64_t n_rows, cudaStream_t s);

}  // namespace vaeassoc
