// libvaeassoc C-ABI implementation: flat parameter layout, op schedule of the train step, CUDA-graph replay,
// NCCL data parallelism.  See include/vaeassoc.h for the contract and the reference call sites each entry
// point replaces.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vaeassoc.h"
#include "common.cuh"
#include "kernels.h"

using namespace vaeassoc;

namespace {

// -------------------------------------------------------------------------------------------------------
// errors
// -------------------------------------------------------------------------------------------------------
thread_local std::string g_create_error;

[[noreturn]] void fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw std::runtime_error(buf);
}

#define CUDA_OK(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

inline int64_t round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// -------------------------------------------------------------------------------------------------------
// NCCL through dlopen (torch ships libnccl.so.2; nothing is linked at build time)
// -------------------------------------------------------------------------------------------------------
struct NcclId { char internal[128]; };
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommGetAsyncError)(void*, int*) = nullptr;
  int (*CommAbort)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  void load(const char* path) {
    if (lib) return;
    lib = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) fail("dlopen(%s) failed: %s", path ? path : "libnccl.so.2", dlerror());
#define NCCL_SYM(field, name)                                    \
    *reinterpret_cast<void**>(&field) = dlsym(lib, name);        \
    if (!field) fail("dlsym(%s) failed", name);
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(AllReduce, "ncclAllReduce")
    NCCL_SYM(Broadcast, "ncclBroadcast")
    NCCL_SYM(CommGetAsyncError, "ncclCommGetAsyncError")
    NCCL_SYM(CommAbort, "ncclCommAbort")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
  }
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;
constexpr int kNcclFloat = 7, kNcclInt64 = 4, kNcclSum = 0;

// -------------------------------------------------------------------------------------------------------
// schedule
// -------------------------------------------------------------------------------------------------------
struct Op {
  std::string name;
  std::function<void(cudaStream_t)> run;
  double flops = 0, bytes = 0;
  int launches = 1;
  int kind = -1;        // tensor-core contraction: 0 NN, 1 NT, 2 TN (else -1) and its final arguments
  GemmArgs gargs;
  bool side = false;   // weight-gradient contraction: nothing downstream in the backward chain reads its result, so it
                       // runs on the modality's side stream, concurrently with the dgrad chain
};

struct Mod {               // one modality (dense variant)
  vaeassoc_modality cfg;
  int ni, nip, r1, r1p, r2, r2p, nz, nh, nhp;
  // float offsets in the flat buffers
  int64_t W1, b1, W2, b2, Wh, bh, V1, c1, V2, c2, Vo, co;
  // activations / gradients (device)
  float *xin[2] = {nullptr, nullptr};   // dense [B, ni] upload buffers (host API, double buffered)
  float *xs = nullptr, *h1 = nullptr, *h2 = nullptr, *hd = nullptr, *z = nullptr, *g1 = nullptr, *g2 = nullptr,
        *xh = nullptr, *da = nullptr, *dg2 = nullptr, *dg1 = nullptr, *dz = nullptr, *dhd = nullptr,
        *dh2 = nullptr, *dh1 = nullptr, *gstat = nullptr, *lat_loss = nullptr, *rec_loss = nullptr,
        *partials = nullptr;
  int recon_blocks = 0;
  // relu sign masks of h1, h2, g1, g2 (one bit per element, [B, ceil(width / 32)] words): written by the tcgen05 forward
  // epilogues, read by the dgrad epilogues in place of the fp32 activation tile
  uint32_t *mh1 = nullptr, *mh2 = nullptr, *mg1 = nullptr, *mg2 = nullptr;
  // ---- hidden_conv=True variant (vae_assoc.py:169-199,249-291; deconv.py) -------------------------------------
  bool conv = false;
  int s0 = 0, s1 = 0, s2 = 0, s3 = 0;        // encoder spatial sizes 28 -> 14 -> 7 -> 3
  int e1 = 0, e2 = 0, e3 = 0;                // encoder depths r1, 2 r1, r2
  int q1 = 0, q2 = 0, q3 = 0;                // decoder depths g1, g1/2, g2
  int t1 = 0, t2 = 0, t3 = 0, t4 = 0;        // decoder spatial sizes 3 -> 7 -> 14 -> 28
  int64_t C1 = 0, C2 = 0, C3 = 0, D1 = 0, d1 = 0, D2 = 0, d2 = 0, D3 = 0, d3 = 0, D4 = 0, d4 = 0;   // flat offsets (Wo/bo reuse Vo/co)
  float *P1 = nullptr, *P2 = nullptr, *P3 = nullptr;             // encoder patch matrices (kept for the wgrads)
  float *l05 = nullptr, *l1 = nullptr, *dl1 = nullptr, *dl05 = nullptr;
  float *cols1 = nullptr, *cols2 = nullptr, *cols3 = nullptr, *cols4 = nullptr;   // deconv columns; reused as backward patches
  float *o1 = nullptr, *o2 = nullptr, *o3 = nullptr, *dd3 = nullptr, *dd2 = nullptr, *dd1 = nullptr;
  // synthetic generator state
  float *P = nullptr, *inv_std = nullptr;
  uint32_t proj_seed_built = 0xFFFFFFFFu;
};

}  // namespace

struct vaeassoc_ctx {
  vaeassoc_config cfg;
  std::mutex mu;
  std::string err;
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr, comm_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  cudaEvent_t ev_bucket = nullptr, ev_comm = nullptr;
  int64_t submit_count = 0;
  static constexpr int kUploadRing = 8;
  cudaEvent_t ev_upload[kUploadRing] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // H2D of submit k done
  std::vector<void*> allocs;
  std::vector<void*> op_allocs;   // per-op workspaces of the current schedule (split partial sums); rebuilt with the ops
  std::vector<Mod> mods;
  std::vector<vaeassoc_tensor_info> tensors;
  int64_t n_flat = 0;          // floats in the parameter part (multiple of 32)
  int64_t bucket_split = 0;    // [0, split) decoders, [split, n_flat) encoders
  float *p = nullptr, *g = nullptr, *m = nullptr, *v = nullptr, *p_tf32 = nullptr;   // g has n_flat + 32 floats
  bool shadow_dirty = true;
  float *eps = nullptr, *eps_in[2] = {nullptr, nullptr}, *lat_partials = nullptr, *scalars = nullptr;
  int64_t* idx_in[2] = {nullptr, nullptr};      // batch row indices of the device-resident data-set path (double buffered)
  int64_t indexed_count = 0;
  float *cost_hist = nullptr, *last_cost = nullptr;
  float* host_cost_ring = nullptr;   // pinned; one async D2H of the step's cost per vaeassoc_submit_host
  static constexpr int kHostRing = 4096;
  int hist_cap = 1 << 16;
  int64_t* step_dev = nullptr;
  int lat_blocks = 0;
  bool round_z = false, round_x = false;
  int64_t launches = 0;
  // schedules
  std::vector<Op> ops_fwd_enc, ops_latent_fwd, ops_fwd_dec, ops_loss, ops_bwd_dec, ops_latent_bwd, ops_bwd_enc;
  std::vector<std::vector<Op>> ops_enc_mod, ops_dec_mod;   // per-modality forward slices (encode / decode)
  std::vector<std::vector<Op>> ops_loss_mod, ops_bwd_dec_mod, ops_bwd_enc_mod;   // per-modality: run concurrently
  cudaStream_t side[VAEASSOC_MAX_MODALITIES - 1] = {nullptr, nullptr, nullptr};   // modality m > 0 runs on side[m-1]
  cudaEvent_t ev_fork = nullptr, ev_join[VAEASSOC_MAX_MODALITIES - 1] = {nullptr, nullptr, nullptr};
  cudaStream_t aux_stream = nullptr;      // small kernels off the critical path (gradient memset, cost finalize)
  cudaEvent_t ev_aux_fork = nullptr, ev_aux_join = nullptr;
  cudaStream_t wstream[VAEASSOC_MAX_MODALITIES] = {nullptr, nullptr, nullptr, nullptr};   // wgrad branch of modality m
  cudaEvent_t ev_wfork[VAEASSOC_MAX_MODALITIES] = {nullptr, nullptr, nullptr, nullptr},
              ev_wjoin[VAEASSOC_MAX_MODALITIES] = {nullptr, nullptr, nullptr, nullptr};
  // tensor-core path: one plan (problems + tile tasks) per handle; gsync = row-block counters, then two words per
  // launch site (queue head, clusters-left), all self-cleaning (zero between launches)
  GroupPlan* gplan = nullptr;
  uint32_t* gsync = nullptr;
  int n_ctr = 0, n_ctr_half = 0, max_sites = 0;
  struct Seg { int site = 0, reset_first = 0, reset_count = 0; };
  Seg seg_enc, seg_dec, seg_bwd_dec, seg_bwd_enc;     // fused segments of the train step (dense modalities, tf32)
  bool fused = false;
  // two-launch form: the latent stages run as elementwise tasks INSIDE the tile kernel, so that encoder + latent +
  // decoder forward are one launch and decoder + latent + encoder backward another (build_segments)
  Seg seg_fwd, seg_bwd;
  bool elt_built = false;
  // one-launch form: forward, reconstruction losses (in the output-layer epilogues), cost finalize and backward as ONE
  // launch of the tile kernel; x_hat / per-row reconstruction losses are then produced on demand (probe_get)
  Seg seg_step;
  bool one_built = false;
  // single-GPU one-launch schedule: Adam hands the gradient buffer back cleared, the train graph has no memset
  bool g_zero = false;                // the gradient buffer holds zeros (stream order)
  // one-launch form: the heads layer (N = 2 n_z) and the decoder input layer's dgrad (N = n_z) run as split-K tasks that
  // add into hd / dz; the staging kernel clears those accumulators
  bool split_heads = false;
  cudaGraphExec_t graph_train_nz = nullptr; int graph_train_nz_nodes = 0;
  unsigned recon_stale = 0;           // bit m: the last step ran the loss-fused form, d.xh / d.rec_loss of modality m are not current
  FinalizeArgs fin_one;
  int lat_blocks_elt = 0;
  bool lat_mode_elt = false;          // which latent kernel wrote lat_partials last (layout of the block partials)
  LatentArgs lat_fwd_args;
  LatentBwdArgs lat_bwd_args;
  bool recon_round_da[VAEASSOC_MAX_MODALITIES] = {false, false, false, false};
  bool masks_in_use[VAEASSOC_MAX_MODALITIES] = {false, false, false, false};   // relu sign masks written / read this schedule
  int dp_single = -1;                                 // VAEASSOC_DP_SINGLE=0/1; default (-1): one all-reduce per step iff fused and world > 4
  bool force_dynamic = false;                         // VAEASSOC_DYNAMIC_FIRST: the data-parallel task-queue mode on one GPU (tests)
  std::vector<Op> ops_colsum_dec, ops_colsum_enc;     // bias gradients that no GEMM epilogue produces (d a, d heads)
  // graphs
  cudaGraphExec_t graph_train = nullptr, graph_grad = nullptr, graph_a1 = nullptr, graph_a2 = nullptr,
                  graph_adam = nullptr;
  int graph_train_nodes = 0, graph_grad_nodes = 0, graph_a1_nodes = 0, graph_a2_nodes = 0, graph_adam_nodes = 0;
  // inference surface with host buffers (vaeassoc_infer_host): per (kind, modality set) one captured graph =
  // H2D of the inputs, the forward launches, packing of the results, ONE D2H
  struct Infer {
    float* pin_in = nullptr;  float* pin_out = nullptr;     // pinned staging, sized for a full batch of every modality
    float* xin[VAEASSOC_MAX_MODALITIES] = {nullptr, nullptr, nullptr, nullptr};   // device: dense [B, n_input_m]
    float* zin = nullptr;     // device: [B, n_z] (z of generate / injected eps of reconstruct)
    float* pack = nullptr;    // device: results of all modalities, contiguous
    cudaGraphExec_t graph[3][1 << VAEASSOC_MAX_MODALITIES][2] = {};   // [kind][modality mask][eps injected]
    int nodes[3][1 << VAEASSOC_MAX_MODALITIES][2] = {};
  } inf;
  // comm
  void* comm = nullptr;
  int rank = 0, world = 1;
  uint64_t comm_calls = 0;
  // peer-memory data-parallel step (peer_adam.cu): the arena of every rank mapped through cudaIpc
  float* arena = nullptr;
  int64_t arena_floats = 0;
  struct Peer {
    bool on = false;
    void* mapped[kMaxPeers] = {};     // what cudaIpcOpenMemHandle returned for rank r (block base)
    float* base[kMaxPeers] = {};      // arena of rank r in THIS process' address space (own rank: c->arena)
    uint32_t* flags = nullptr;        // local arrival words (inside the arena)
    cudaGraphExec_t graph = nullptr;  // a1 + a2 + peer_adam: the whole data-parallel step as one graph
    int graph_nodes = 0;
    bool slots_stale = false;         // m / v of the shards owned by peers are behind (refreshed on demand)
    bool pending = false;             // the last step's kernel left without waiting for the peers' parameter stores:
                                      // the next step's graph (or peer_quiesce) waits for them first
    float* mc = nullptr;              // NVLS multicast address of the arena (symmetric attach), or null
    bool symmetric = false;           // the arena lives in caller-provided symmetric memory (not cudaIpc-mapped)
    unsigned long long* tl = nullptr; // VAEASSOC_PEER_TIMELINE: phase stamps summed on the device, printed at detach
  } peer;

  // Own bounds check (compute-sanitizer is closed on the GPU pool): every device buffer is followed by a 256-byte guard
  // filled with kGuardByte; nothing in the library may WRITE there (the 3-D tensor maps of MN-major GEMM operands may
  // READ up to 124 B past the last row, gemm_group.cu -- those lanes only feed clipped outputs).
  // vaeassoc_debug_guard_check compares every guard with the pattern.
  static constexpr int kGuardBytes = 256;
  static constexpr int kGuardByte = 0x2b;        // 0x2b2b2b2b as a float is 6.1e-13: harmless where a guard is read
  std::vector<std::pair<char*, size_t>> guards;
  void add_guard(void* addr, size_t len) {
    CUDA_OK(cudaMemset(addr, kGuardByte, len));
    guards.emplace_back(reinterpret_cast<char*>(addr), len);
  }
  template <typename T>
  T* dalloc(int64_t n, bool zero = true) {
    void* ptr = nullptr;
    const size_t payload = (size_t)std::max<int64_t>(n, 1) * sizeof(T);
    const size_t bytes = payload + kGuardBytes;
    CUDA_OK(cudaMalloc(&ptr, bytes));
    if (zero) CUDA_OK(cudaMemset(ptr, 0, bytes));
    allocs.push_back(ptr);
    add_guard(reinterpret_cast<char*>(ptr) + payload, kGuardBytes);
    return reinterpret_cast<T*>(ptr);
  }
  // zero-initialised workspace owned by one op of the schedule: ops of different streams never share one
  float* op_ws(int64_t floats) {
    if (floats <= 0) return nullptr;
    void* ptr = nullptr;
    CUDA_OK(cudaMalloc(&ptr, (size_t)floats * sizeof(float) + kGuardBytes));
    CUDA_OK(cudaMemset(ptr, 0, (size_t)floats * sizeof(float)));
    op_allocs.push_back(ptr);
    char* gaddr = reinterpret_cast<char*>(ptr) + (size_t)floats * sizeof(float);
    CUDA_OK(cudaMemset(gaddr, kGuardByte, kGuardBytes));
    op_guards.emplace_back(gaddr, (size_t)kGuardBytes);
    return reinterpret_cast<float*>(ptr);
  }
  std::vector<std::pair<char*, size_t>> op_guards;   // guards of the per-op workspaces (rebuilt with the ops)
  void free_op_ws() {
    for (void* p : op_allocs) cudaFree(p);
    op_guards.clear();
    op_allocs.clear();
  }
};

namespace {

using Ctx = vaeassoc_ctx;
void build_ops(Ctx* c);
FinalizeArgs finalize_args(Ctx* c, int advance);
void peer_quiesce(Ctx* c);
void peer_detach(Ctx* c);
void refresh_peer_slots(Ctx* c);

// the task queue hands out EVERY task (first one included) whenever a kernel that waits on a peer GPU may hold SMs next to
// the tile kernel: the NCCL schedules.  The peer-memory step runs alone on the stream, so it keeps the static first task.
bool dyn_first(const Ctx* c) { return c->force_dynamic || (c->comm != nullptr && !c->peer.on); }

int64_t global_batch(const Ctx* c) { return c->cfg.global_batch > 0 ? c->cfg.global_batch : c->cfg.batch_size; }
int fwd_act(const Ctx* c) { return c->cfg.transfer_fct == VAEASSOC_SOFTPLUS ? ACT_SOFTPLUS : ACT_RELU; }

// ---- layout --------------------------------------------------------------------------------------------
void add_tensor(Ctx* c, int m, const char* scope, int scope_entry, int var_idx, const char* role, int ndim,
                std::initializer_list<int> shape, int64_t offset, int64_t rows, int64_t cols, int64_t ld) {
  vaeassoc_tensor_info t;
  memset(&t, 0, sizeof t);
  char sc[40];
  if (scope_entry == 0) snprintf(sc, sizeof sc, "%s", scope);
  else snprintf(sc, sizeof sc, "%s_%d", scope, scope_entry);
  if (var_idx == 0) snprintf(t.name, sizeof t.name, "%s/Variable", sc);
  else snprintf(t.name, sizeof t.name, "%s/Variable_%d", sc, var_idx);
  snprintf(t.role, sizeof t.role, "%s", role);
  t.modality = m;
  t.ndim = ndim;
  int i = 0;
  for (int s : shape) t.shape[i++] = s;
  t.offset = offset; t.rows = rows; t.cols = cols; t.ld = ld;
  c->tensors.push_back(t);
}

void add_named(Ctx* c, int m, const char* name, const char* role, int ndim, std::initializer_list<int> shape,
               int64_t offset, int64_t rows, int64_t cols, int64_t ld) {
  vaeassoc_tensor_info t;
  memset(&t, 0, sizeof t);
  snprintf(t.name, sizeof t.name, "%s", name);
  snprintf(t.role, sizeof t.role, "%s", role);
  t.modality = m; t.ndim = ndim;
  int i = 0;
  for (int s : shape) t.shape[i++] = s;
  t.offset = offset; t.rows = rows; t.cols = cols; t.ld = ld;
  c->tensors.push_back(t);
}

// TensorFlow SAME rule for stride-2, VALID otherwise (vae_assoc.py:174-197)
inline int same_out(int n, int s) { return (n + s - 1) / s; }
inline int same_pad_before(int n, int k, int s) { int o = same_out(n, s); int t = (o - 1) * s + k - n; return t > 0 ? t / 2 : 0; }

void build_layout(Ctx* c) {
  const int M = c->cfg.n_modalities;
  const int nz = c->cfg.n_z;
  int64_t off = 0;
  auto take = [&](int64_t rows, int64_t ld) { int64_t o = off; off = round_up(off + rows * ld, 32); return o; };
  c->mods.resize(M);
  for (int m = 0; m < M; ++m) {
    Mod& d = c->mods[m];
    d.cfg = c->cfg.mod[m];
    d.conv = d.cfg.hidden_conv != 0;
    // row pitches: dense modalities round to 32 floats so that every row starts on a 128-byte line (one L2 line per
    // 32-float TMA box row / epilogue row segment instead of two); the conv modality keeps dense NHWC images (4 floats)
    int al = d.conv ? 4 : 32;
    if (const char* e = getenv("VAEASSOC_PITCH_ALIGN")) { const int v = atoi(e); if (!d.conv && v >= 4 && v % 4 == 0) al = v; }
    d.ni = d.cfg.n_input; d.nip = (int)round_up(d.ni, al);
    d.r1 = d.cfg.n_hidden_recog_1; d.r1p = (int)round_up(d.r1, al);
    d.r2 = d.cfg.n_hidden_recog_2; d.r2p = (int)round_up(d.r2, al);
    d.nz = nz; d.nh = 2 * nz; d.nhp = (int)round_up(d.nh, 4);
    if (d.ni <= 0 || d.r1 <= 0 || d.r2 <= 0) fail("modality %d: layer sizes must be positive", m);
    if (d.conv) {
      // encoder 28 -> 14 -> 7 (5x5, stride 2, SAME) -> 3 (5x5 VALID); decoder 1 -> 3 (3x3) -> 7 (5x5) -> 14 -> 28
      if (!d.cfg.binary)
        fail("modality %d: hidden_conv with a Gaussian output is ill-formed in the reference (vae_assoc.py:299-303 "
             "multiplies [B, n_input] by [n_hidden_recog_2, n_input])", m);
      d.s0 = (int)lround(sqrt((double)d.ni));
      if (d.s0 * d.s0 != d.ni) fail("modality %d: hidden_conv needs a square image, n_input = %d", m, d.ni);
      d.s1 = same_out(d.s0, 2); d.s2 = same_out(d.s1, 2); d.s3 = d.s2 - 5 + 1;
      if (d.s3 < 1) fail("modality %d: image too small for the conv encoder", m);
      d.e1 = d.r1; d.e2 = 2 * d.r1; d.e3 = d.r2;
      d.q1 = d.cfg.n_hidden_gener_1; d.q2 = d.q1 / 2; d.q3 = d.cfg.n_hidden_gener_2;
      if (d.q1 < 2 || d.q3 < 1) fail("modality %d: n_hidden_gener_1 >= 2 and n_hidden_gener_2 >= 1 required", m);
      d.t1 = 3; d.t2 = d.t1 + 4; d.t3 = d.t2 * 2; d.t4 = d.t3 * 2;
      if (d.t4 * d.t4 != d.ni)
        fail("modality %d: the deconv chain of vae_assoc.py:251-277 produces %dx%d, n_input = %d", m, d.t4, d.t4, d.ni);
      if ((d.e1 % 4) || (d.e2 % 4) || (d.e3 % 4) || (d.q1 % 4) || (d.q2 % 4) || (d.q3 % 4))
        fail("modality %d: conv depths must be multiples of 4 (16-byte rows)", m);
      // the flattened conv output plays the role of the second hidden layer for the heads
      d.r2 = d.s3 * d.s3 * d.e3; d.r2p = d.r2;
    }
  }
  // bucket 0: decoders, in the order their gradients complete (output layer first)
  for (int m = 0; m < M; ++m) {
    Mod& d = c->mods[m];
    if (d.conv) {
      const int nzp = (int)round_up(nz, 4);
      d.Vo = take(d.ni, d.nip); d.co = take(1, d.nip);                       // Wo [n_input, n_input], bo   (:287-288)
      d.D4 = take(25 * 1, d.q3);      d.d4 = take(1, 4);                     // [5,5,1,g2]      (:273-277)
      d.D3 = take(25 * d.q3, d.q2);   d.d3 = take(1, d.q3);                  // [5,5,g2,g1/2]   (:268-272)
      d.D2 = take(25 * d.q2, d.q1);   d.d2 = take(1, d.q2);                  // [5,5,g1/2,g1]   (:263-267)
      d.D1 = take(9 * d.q1, nzp);     d.d1 = take(1, d.q1);                  // [3,3,g1,n_z]    (:251-255)
      continue;
    }
    d.Vo = take(d.r2, d.nip); d.co = take(1, d.nip);
    d.V2 = take(d.r1, d.r2p); d.c2 = take(1, d.r2p);
    d.V1 = take(d.nz, d.r1p); d.c1 = take(1, d.r1p);
  }
  c->bucket_split = off;
  // bucket 1: encoders
  for (int m = 0; m < M; ++m) {
    Mod& d = c->mods[m];
    d.Wh = take(d.r2, d.nhp); d.bh = take(1, d.nhp);
    if (d.conv) {
      d.C3 = take(25 * d.e2, d.e3); d.C2 = take(25 * d.e1, d.e2); d.C1 = take(25, d.e1);   // (:194-197,:179-183,:174-178)
      continue;
    }
    d.W2 = take(d.r1, d.r2p); d.b2 = take(1, d.r2p);
    d.W1 = take(d.ni, d.r1p); d.b1 = take(1, d.r1p);
  }
  c->n_flat = off;
  // logical tensors in the reference's tf.Variable creation order (vae_assoc.py:93-111: per modality
  // recognition (8 variables, scope "<scope>") then generator (6 variables, name scope "<scope>_1"))
  // default scope names follow the reference script (vae_assoc_ujichar_img_jnt.py:54,64)
  // scope = na["scope"] (vae_assoc.py:168,248); default names follow the reference script (vae_assoc_ujichar_img_jnt.py:54,64)
  static const char* default_scopes[4] = {"image", "joint", "modal2", "modal3"};
  for (int m = 0; m < M; ++m) {
    const Mod& d = c->mods[m];
    char scbuf[33];
    memcpy(scbuf, d.cfg.scope, 32); scbuf[32] = 0;
    const char* sc = scbuf[0] ? scbuf : default_scopes[m];
    for (int k = 0; k < m; ++k) {
      char other[33];
      memcpy(other, c->mods[k].cfg.scope, 32); other[32] = 0;
      if (!strcmp(sc, other[0] ? other : default_scopes[k])) fail("modalities %d and %d share the variable scope '%s'", k, m, sc);
    }
    if (d.conv) {
      // conv_2d weights are plain tf.Variables (vae_assoc.py:471-473); deconv2d ones are prettytensor variables
      // 'weights' / 'bias' of layer scopes deconv2d, deconv2d_1, ... (deconv.py:92,110-114)
      char nm[48];
      const int nzp = (int)round_up(nz, 4);
      add_tensor(c, m, sc, 0, 0, "C1", 4, {5, 5, 1, d.e1}, d.C1, 25, d.e1, d.e1);
      add_tensor(c, m, sc, 0, 1, "C2", 4, {5, 5, d.e1, d.e2}, d.C2, 25 * d.e1, d.e2, d.e2);
      add_tensor(c, m, sc, 0, 2, "C3", 4, {5, 5, d.e2, d.e3}, d.C3, 25 * d.e2, d.e3, d.e3);
      add_tensor(c, m, sc, 0, 3, "Wmu", 2, {d.r2, nz}, d.Wh, d.r2, nz, d.nhp);
      add_tensor(c, m, sc, 0, 4, "bmu", 1, {nz}, d.bh, 1, nz, d.nhp);
      add_tensor(c, m, sc, 0, 5, "Wls", 2, {d.r2, nz}, d.Wh + nz, d.r2, nz, d.nhp);
      add_tensor(c, m, sc, 0, 6, "bls", 1, {nz}, d.bh + nz, 1, nz, d.nhp);
      const int64_t woff[4] = {d.D1, d.D2, d.D3, d.D4}, boff[4] = {d.d1, d.d2, d.d3, d.d4};
      const int kk[4] = {3, 5, 5, 5}, od[4] = {d.q1, d.q2, d.q3, 1}, id[4] = {nz, d.q1, d.q2, d.q3};
      const int64_t wld[4] = {nzp, d.q1, d.q2, d.q3};
      const char* wr[4] = {"D1", "D2", "D3", "D4"}; const char* br[4] = {"d1", "d2", "d3", "d4"};
      for (int l = 0; l < 4; ++l) {
        if (l == 0) snprintf(nm, sizeof nm, "%s_1/deconv2d/weights", sc); else snprintf(nm, sizeof nm, "%s_1/deconv2d_%d/weights", sc, l);
        add_named(c, m, nm, wr[l], 4, {kk[l], kk[l], od[l], id[l]}, woff[l], (int64_t)kk[l] * kk[l] * od[l], id[l], wld[l]);
        if (l == 0) snprintf(nm, sizeof nm, "%s_1/deconv2d/bias", sc); else snprintf(nm, sizeof nm, "%s_1/deconv2d_%d/bias", sc, l);
        add_named(c, m, nm, br[l], 1, {od[l]}, boff[l], 1, od[l], (int64_t)round_up(od[l], 4));
      }
      add_tensor(c, m, sc, 1, 0, "Wo", 2, {d.ni, d.ni}, d.Vo, d.ni, d.ni, d.nip);
      add_tensor(c, m, sc, 1, 1, "bo", 1, {d.ni}, d.co, 1, d.ni, d.nip);
      continue;
    }
    add_tensor(c, m, sc, 0, 0, "W1", 2, {d.ni, d.r1}, d.W1, d.ni, d.r1, d.r1p);
    add_tensor(c, m, sc, 0, 1, "b1", 1, {d.r1}, d.b1, 1, d.r1, d.r1p);
    add_tensor(c, m, sc, 0, 2, "W2", 2, {d.r1, d.r2}, d.W2, d.r1, d.r2, d.r2p);
    add_tensor(c, m, sc, 0, 3, "b2", 1, {d.r2}, d.b2, 1, d.r2, d.r2p);
    add_tensor(c, m, sc, 0, 4, "Wmu", 2, {d.r2, nz}, d.Wh, d.r2, nz, d.nhp);
    add_tensor(c, m, sc, 0, 5, "bmu", 1, {nz}, d.bh, 1, nz, d.nhp);
    add_tensor(c, m, sc, 0, 6, "Wls", 2, {d.r2, nz}, d.Wh + nz, d.r2, nz, d.nhp);
    add_tensor(c, m, sc, 0, 7, "bls", 1, {nz}, d.bh + nz, 1, nz, d.nhp);
    add_tensor(c, m, sc, 1, 0, "V1", 2, {nz, d.r1}, d.V1, nz, d.r1, d.r1p);
    add_tensor(c, m, sc, 1, 1, "c1", 1, {d.r1}, d.c1, 1, d.r1, d.r1p);
    add_tensor(c, m, sc, 1, 2, "V2", 2, {d.r1, d.r2}, d.V2, d.r1, d.r2, d.r2p);
    add_tensor(c, m, sc, 1, 3, "c2", 1, {d.r2}, d.c2, 1, d.r2, d.r2p);
    add_tensor(c, m, sc, 1, 4, "Vo", 2, {d.r2, d.ni}, d.Vo, d.r2, d.ni, d.nip);
    add_tensor(c, m, sc, 1, 5, "co", 1, {d.ni}, d.co, 1, d.ni, d.nip);
  }
}

void alloc_buffers(Ctx* c) {
  const int64_t B = c->cfg.batch_size;
  const int nz = c->cfg.n_z;
  {
    // ONE allocation ("arena") holds the flat buffers and the peer arrival words, so that a single cudaIpc handle maps
    // everything a peer rank touches in the data-parallel step (peer_adam.cu).  Every sub-buffer starts on a 256-byte
    // boundary and is followed by >= 256 bytes of slack (the 3-D tensor maps of MN-major operands may read 124 B past
    // the last row, gemm_group.cu).
    const int64_t slot = round_up(c->n_flat + 32, 64) + 64;     // floats per flat buffer incl. slack
    c->arena_floats = 5 * slot + 64;
    c->arena = c->dalloc<float>(c->arena_floats);
    c->p = c->arena;
    c->g = c->arena + slot;
    c->m = c->arena + 2 * slot;
    c->v = c->arena + 3 * slot;
    c->p_tf32 = c->arena + 4 * slot;
    c->peer.flags = reinterpret_cast<uint32_t*>(c->arena + 5 * slot);   // 2 x kMaxPeers arrival words, then sync[2]
    for (int i = 0; i < 5; ++i) {           // the slack behind each flat buffer doubles as its guard
      const int64_t used = (i == 1) ? c->n_flat + 32 : c->n_flat;
      c->add_guard(c->arena + i * slot + used, (size_t)(slot - used) * sizeof(float));
    }
  }
  {
    // row-block counters: 12 activation / gradient tensors per modality x row blocks of 256; then the launch sites
    const int64_t rb = (B + 255) / 256;
    c->n_ctr_half = (int)(c->cfg.n_modalities * 13 * rb);    // full counters, then the half-tile counters of the same tensors
    c->n_ctr = 2 * c->n_ctr_half;
    c->max_sites = 1024;
    c->gsync = c->dalloc<uint32_t>(c->n_ctr + 2 * c->max_sites);
  }
  c->eps = c->dalloc<float>(B * nz);
  c->eps_in[0] = c->dalloc<float>(B * nz);
  c->eps_in[1] = c->dalloc<float>(B * nz);
  c->idx_in[0] = c->dalloc<int64_t>(B);
  c->idx_in[1] = c->dalloc<int64_t>(B);
  c->lat_partials = c->dalloc<float>((int64_t)kMaxPartialBlocks * kCostSlots);
  c->scalars = c->dalloc<float>(16);
  c->cost_hist = c->dalloc<float>(c->hist_cap);
  c->last_cost = c->dalloc<float>(1);
  c->step_dev = c->dalloc<int64_t>(1);
  CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&c->host_cost_ring), sizeof(float) * Ctx::kHostRing));
  {
    int64_t width = 0;
    for (int m = 0; m < c->cfg.n_modalities; ++m) {
      c->inf.xin[m] = c->dalloc<float>(B * c->mods[m].ni);
      width += c->mods[m].ni;
    }
    c->inf.zin = c->dalloc<float>(B * nz);
    c->inf.pack = c->dalloc<float>(B * width);
    CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&c->inf.pin_in), sizeof(float) * (size_t)(B * (width + nz))));
    CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&c->inf.pin_out), sizeof(float) * (size_t)(B * width)));
  }
  for (Mod& d : c->mods) {
    d.xin[0] = c->dalloc<float>(B * d.ni);
    d.xin[1] = c->dalloc<float>(B * d.ni);
    d.xs = c->dalloc<float>(B * d.nip);
    d.h1 = c->dalloc<float>(B * d.r1p);  d.h2 = c->dalloc<float>(B * d.r2p);
    d.hd = c->dalloc<float>(B * d.nh);   d.z = c->dalloc<float>(B * nz);
    d.g1 = c->dalloc<float>(B * d.r1p);  d.g2 = c->dalloc<float>(B * std::max(d.r2p, d.nip));   // conv: g2 = last deconv output [B, n_input]
    d.xh = c->dalloc<float>(B * d.nip);  d.da = c->dalloc<float>(B * d.nip);
    d.dg2 = c->dalloc<float>(B * std::max(d.r2p, d.nip)); d.dg1 = c->dalloc<float>(B * d.r1p);
    d.dz = c->dalloc<float>(B * nz);     d.dhd = c->dalloc<float>(B * d.nh);
    d.dh2 = c->dalloc<float>(B * d.r2p); d.dh1 = c->dalloc<float>(B * d.r1p);
    d.gstat = c->dalloc<float>(B * d.nh);
    d.lat_loss = c->dalloc<float>(B);    d.rec_loss = c->dalloc<float>(B);
    d.partials = c->dalloc<float>((int64_t)kMaxPartialBlocks * kCostSlots);
    d.P = c->dalloc<float>(4 * d.ni);    d.inv_std = c->dalloc<float>(d.ni);
    if (!d.conv) {
      d.mh1 = c->dalloc<uint32_t>(B * ((d.r1 + 31) / 32)); d.mg1 = c->dalloc<uint32_t>(B * ((d.r1 + 31) / 32));
      d.mh2 = c->dalloc<uint32_t>(B * ((d.r2 + 31) / 32)); d.mg2 = c->dalloc<uint32_t>(B * ((d.r2 + 31) / 32));
    }
    if (d.conv) {
      const int64_t n1 = B * d.s1 * d.s1, n2 = B * d.s2 * d.s2, n3 = B * d.s3 * d.s3;       // encoder rows
      const int64_t m1 = B * d.t1 * d.t1, m2 = B * d.t2 * d.t2, m3 = B * d.t3 * d.t3;       // decoder rows (per input pixel)
      d.P1 = c->dalloc<float>(n1 * 28);             d.l05 = c->dalloc<float>(n1 * d.e1);  d.dl05 = c->dalloc<float>(n1 * d.e1);
      d.P2 = c->dalloc<float>(n2 * 25 * d.e1);      d.l1 = c->dalloc<float>(n2 * d.e2);   d.dl1 = c->dalloc<float>(n2 * d.e2);
      d.P3 = c->dalloc<float>(n3 * 25 * d.e2);
      d.cols1 = c->dalloc<float>(B * 9 * d.q1);     d.o1 = c->dalloc<float>(m1 * d.q1);   d.dd1 = c->dalloc<float>(m1 * d.q1);
      d.cols2 = c->dalloc<float>(std::max(m1 * 25 * d.q2, n3 * 25 * d.e2));
      d.o2 = c->dalloc<float>(m2 * d.q2);           d.dd2 = c->dalloc<float>(m2 * d.q2);
      d.cols3 = c->dalloc<float>(std::max(m2 * 25 * d.q3, n2 * 25 * d.e1));
      d.o3 = c->dalloc<float>(m3 * d.q3);           d.dd3 = c->dalloc<float>(m3 * d.q3);
      d.cols4 = c->dalloc<float>(m3 * 28);
    }
  }
}

// ---- op construction -------------------------------------------------------------------------------------
enum { KIND_NN = 0, KIND_NT = 1, KIND_TN = 2 };

// w_off >= 0: operand B is the weight at that flat offset (master fp32 for SIMT, tf32-rounded shadow for tcgen05)
// round_out: the stored result feeds a tcgen05 kind::tf32 GEMM and is rounded (RNA) by this producer
Op make_gemm(Ctx* c, const char* name, int m, int kind, GemmArgs a, int64_t w_off, bool round_out) {
  Op op;
  op.name = std::string(name) + "." + std::to_string(m);
  op.side = (kind == KIND_TN);
  const bool tf32 = c->cfg.precision == VAEASSOC_TF32;
  a.round_out = (tf32 && round_out) ? 1 : 0;
  op.flops = 2.0 * a.M * a.N * a.K;
  op.bytes = 4.0 * ((double)a.M * a.K + (double)a.K * a.N + (double)a.M * a.N * (kind == KIND_NT && a.aux ? 2 : 1));
  if (w_off >= 0) a.B = c->p_tf32 + w_off;
  if (tf32 && tc_supported(kind, a)) {
    // one launch of the persistent tile kernel over this contraction's tiles (no dependencies); the fused train-step
    // segments (build_segments) reuse the same arguments
    char err[256] = {0};
    const int site = group_begin(c->gplan);
    if (site >= c->max_sites - 1) fail("too many tensor-core launch sites");
    const int prob = group_add_problem(c->gplan, kind, a, err, sizeof err);
    if (prob < 0) fail("tcgen05 plan for %s failed: %s", op.name.c_str(), err);
    const int tm = group_problem_tiles_m(c->gplan, prob), tn = group_problem_tiles_n(c->gplan, prob);
    const int kb = group_problem_kblocks(c->gplan, prob);
    if (kind == KIND_TN) {
      // split the batch contraction into row-block ranges so that every SM pair gets a task
      const int kpr = group_problem_kb_per_rowblock(c->gplan, prob);
      const int rbs = (kb + kpr - 1) / kpr;
      int splits = std::max(1, std::min(rbs, (kNumSMs / 2) / std::max(1, tm * tn)));
      if (a.splitk > 1) splits = std::min(a.splitk, rbs);
      const int per = (rbs + splits - 1) / splits;
      for (int r0 = 0; r0 < rbs; r0 += per)
        for (int i = 0; i < tm; ++i)
          for (int j = 0; j < tn; ++j)
            group_add_task(c->gplan, prob, i, j, r0 * kpr, std::min(per * kpr, kb - r0 * kpr), -1, 0, 0, -1, 0, -1);
    } else {
      for (int i = 0; i < tm; ++i)
        for (int j = 0; j < tn; ++j) group_add_task(c->gplan, prob, i, j, 0, kb, -1, 0, 0, -1, 0, -1);
    }
    if (!group_end(c->gplan, err, sizeof err)) fail("%s", err);
    op.name += ".tc";
    op.kind = kind;
    op.gargs = a;
    Ctx* cc = c;
    if (kind == KIND_TN && a.bias_grad) {
      const GemmArgs b = a;
      op.launches = 2;
      float* cws = c->op_ws(colsum_ws_floats(b.K, b.N));
      op.run = [cc, site, b, cws](cudaStream_t s) {
        group_launch(cc->gplan, site, cc->gsync + cc->n_ctr + 2 * site, 0, 0, dyn_first(cc), s);
        launch_colsum(b.B, b.ldb, b.K, b.N, b.bias_grad, cws, s);
      };
    } else {
      op.run = [cc, site](cudaStream_t s) {
        group_launch(cc->gplan, site, cc->gsync + cc->n_ctr + 2 * site, 0, 0, dyn_first(cc), s);
      };
    }
    return op;
  }
  if (w_off >= 0) a.B = c->p + w_off;
  if (skinny_supported(kind, a)) {
    op.name += ".sk";
    a.ws = c->op_ws(gemm_skinny_ws_floats(kind, a));
    if (kind == KIND_TN) op.launches = 2;
    op.run = [kind, a](cudaStream_t s) { launch_gemm_skinny(kind, a, s); };
    return op;
  }
  switch (kind) {
    case KIND_NN: op.run = [a](cudaStream_t s) { launch_gemm_nn_simt(a, s); }; break;
    case KIND_NT: op.run = [a](cudaStream_t s) { launch_gemm_nt_simt(a, s); }; break;
    default:
      a.ws = c->op_ws(gemm_tn_simt_ws_floats(a));
      if (a.ws) op.launches = 2;
      op.run = [a](cudaStream_t s) { launch_gemm_tn_simt(a, s); };
      break;
  }
  return op;
}

GemmArgs gemm_fwd(int B, int N, int K, const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                  float* C, int64_t ldc, int act) {
  GemmArgs a;
  a.M = B; a.N = N; a.K = K; a.A = A; a.lda = lda; a.B = W; a.ldb = ldw; a.C = C; a.ldc = ldc; a.bias = bias; a.act = act;
  return a;
}
// dX[B,K] = (dY[B,N] . W[K,N]^T) (*) act'(H[B,K])
GemmArgs gemm_dgrad(int B, int K, int N, const float* dY, int64_t lddy, const float* W, int64_t ldw, float* dX,
                    int64_t lddx, const float* H, int64_t ldh, int act) {
  GemmArgs a;
  a.M = B; a.N = K; a.K = N; a.A = dY; a.lda = lddy; a.B = W; a.ldb = ldw; a.C = dX; a.ldc = lddx;
  a.aux = H; a.ldaux = ldh; a.act = H ? act : ACT_NONE;
  return a;
}
// dW[K,N] += X[B,K]^T . dY[B,N] ; db[N] += colsum(dY)
GemmArgs gemm_wgrad(int B, int K, int N, const float* X, int64_t ldx, const float* dY, int64_t lddy, float* dW,
                    int64_t lddw, float* db) {
  GemmArgs a;
  a.M = K; a.N = N; a.K = B; a.A = X; a.lda = ldx; a.B = dY; a.ldb = lddy; a.C = dW; a.ldc = lddw; a.bias_grad = db;
  return a;
}

void destroy_graphs(Ctx* c) {
  for (cudaGraphExec_t* g : {&c->graph_train, &c->graph_train_nz, &c->graph_grad, &c->graph_a1, &c->graph_a2, &c->graph_adam, &c->peer.graph}) {
    if (*g) cudaGraphExecDestroy(*g);
    *g = nullptr;
  }
  for (auto& k : c->inf.graph)
    for (auto& mk : k)
      for (cudaGraphExec_t& g : mk) {
        if (g) cudaGraphExecDestroy(g);
        g = nullptr;
      }
}

// ---- hidden_conv=True modality: conv / deconv layers as im2col -> GEMM -> col2im (conv.cu) ------------------------
Op make_im2col(const char* name, int m, const float* x, int B, int H, int C, int k, int s, int pb, int OH, float* out,
               int64_t ldo) {
  Op op; op.name = std::string(name) + "." + std::to_string(m);
  Im2colArgs a;
  a.x = x; a.B = B; a.H = H; a.W = H; a.C = C; a.k = k; a.s = s; a.pb = pb; a.OH = OH; a.OW = OH; a.out = out; a.ldo = ldo;
  op.bytes = 4.0 * ((double)B * H * H * C + (double)B * OH * OH * k * k * C);
  op.run = [a](cudaStream_t st) { launch_im2col(a, st); };
  return op;
}
Op make_col2im(const char* name, int m, const float* cols, int64_t ldc, int B, int H, int C, int h, int k, int s, int pb,
               const float* bias, int act, bool round_out, float* out) {
  Op op; op.name = std::string(name) + "." + std::to_string(m);
  Col2imArgs a;
  a.cols = cols; a.ldc = ldc; a.B = B; a.H = H; a.W = H; a.C = C; a.h = h; a.w = h; a.k = k; a.s = s; a.pb = pb;
  a.bias = bias; a.act = act; a.round_out = round_out ? 1 : 0; a.out = out;
  op.bytes = 4.0 * ((double)B * H * H * C + (double)B * h * h * k * k * C);
  op.run = [a](cudaStream_t st) { launch_col2im(a, st); };
  return op;
}
Op make_colsum(Ctx* c, const char* name, int m, const float* X, int64_t ld, int64_t rows, int cols, float* out) {
  Op op; op.name = std::string(name) + "." + std::to_string(m);
  op.bytes = 4.0 * rows * cols;
  float* ws = c->op_ws(colsum_ws_floats(rows, cols));
  op.run = [=](cudaStream_t st) { launch_colsum(X, ld, rows, cols, out, ws, st); };
  return op;
}
GemmArgs gemm_plain(int M, int N, int K, const float* A, int64_t lda, int64_t ldb, float* C, int64_t ldc) {
  GemmArgs a;
  a.M = M; a.N = N; a.K = K; a.A = A; a.lda = lda; a.ldb = ldb; a.C = C; a.ldc = ldc;
  return a;
}

void build_ops_conv(Ctx* c, int m) {
  Mod& d = c->mods[m];
  const int B = c->cfg.batch_size, nz = c->cfg.n_z;
  const bool tf32 = c->cfg.precision == VAEASSOC_TF32;
  const bool R = tf32;                       // conv modality, tf32 mode: every GEMM operand is rounded by its producer
  const int nzp = (int)round_up(nz, 4);
  float* P = c->p; float* G = c->g;
  const int n1 = B * d.s1 * d.s1, n2 = B * d.s2 * d.s2, n3 = B * d.s3 * d.s3;
  const int m1 = B * d.t1 * d.t1, m2 = B * d.t2 * d.t2, m3 = B * d.t3 * d.t3;
  const int pb_s = same_pad_before(d.s0, 5, 2);          // = 1 for 28 -> 14 and 14 -> 7, and for the SAME deconvs
  auto& enc = c->ops_enc_mod[m];
  auto& dec = c->ops_dec_mod[m];
  auto& bd = c->ops_bwd_dec_mod[m];
  auto& be = c->ops_bwd_enc_mod[m];
  auto W = [&](GemmArgs a, int64_t off) { (void)off; return a; };
  (void)W;

  // ---------------- encoder: three bias-free linear convs (vae_assoc.py:172-199), then the heads --------------------
  enc.push_back(make_im2col("enc_im2col1", m, d.xs, B, d.s0, 1, 5, 2, pb_s, d.s1, d.P1, 28));
  enc.push_back(make_gemm(c, "enc_conv1", m, KIND_NN, gemm_plain(n1, d.e1, 25, d.P1, 28, d.e1, d.l05, d.e1), d.C1, R));
  enc.push_back(make_im2col("enc_im2col2", m, d.l05, B, d.s1, d.e1, 5, 2, same_pad_before(d.s1, 5, 2), d.s2, d.P2, 25 * d.e1));
  enc.push_back(make_gemm(c, "enc_conv2", m, KIND_NN, gemm_plain(n2, d.e2, 25 * d.e1, d.P2, 25 * d.e1, d.e2, d.l1, d.e2), d.C2, R));
  enc.push_back(make_im2col("enc_im2col3", m, d.l1, B, d.s2, d.e2, 5, 1, 0, d.s3, d.P3, 25 * d.e2));
  enc.push_back(make_gemm(c, "enc_conv3", m, KIND_NN, gemm_plain(n3, d.e3, 25 * d.e2, d.P3, 25 * d.e2, d.e3, d.h2, d.e3), d.C3, R));
  enc.push_back(make_gemm(c, "fwd_heads", m, KIND_NN,
                          gemm_fwd(B, d.nh, d.r2, d.h2, d.r2p, nullptr, d.nhp, P + d.bh, d.hd, d.nh, ACT_NONE), d.Wh, false));

  // ---------------- decoder: four deconv2d + bias + SIGMOID (deconv_2d default, vae_assoc.py:491), dense 784x784 ----
  dec.push_back(make_gemm(c, "dec_cols1", m, KIND_NT, gemm_plain(B, 9 * d.q1, nz, d.z, nz, nzp, d.cols1, 9 * d.q1), d.D1, false));
  dec.push_back(make_col2im("dec_deconv1", m, d.cols1, 9 * d.q1, B, d.t1, d.q1, 1, 3, 1, 0, P + d.d1, ACT_SIGMOID, R, d.o1));
  dec.push_back(make_gemm(c, "dec_cols2", m, KIND_NT, gemm_plain(m1, 25 * d.q2, d.q1, d.o1, d.q1, d.q1, d.cols2, 25 * d.q2), d.D2, false));
  dec.push_back(make_col2im("dec_deconv2", m, d.cols2, 25 * d.q2, B, d.t2, d.q2, d.t1, 5, 1, 0, P + d.d2, ACT_SIGMOID, R, d.o2));
  dec.push_back(make_gemm(c, "dec_cols3", m, KIND_NT, gemm_plain(m2, 25 * d.q3, d.q2, d.o2, d.q2, d.q2, d.cols3, 25 * d.q3), d.D3, false));
  dec.push_back(make_col2im("dec_deconv3", m, d.cols3, 25 * d.q3, B, d.t3, d.q3, d.t2, 5, 2, pb_s, P + d.d3, ACT_SIGMOID, R, d.o3));
  dec.push_back(make_gemm(c, "dec_cols4", m, KIND_NT, gemm_plain(m3, 25, d.q3, d.o3, d.q3, d.q3, d.cols4, 28), d.D4, false));
  dec.push_back(make_col2im("dec_deconv4", m, d.cols4, 28, B, d.t4, 1, d.t3, 5, 2, pb_s, P + d.d4, ACT_SIGMOID, R, d.g2));
  GemmArgs f_o = gemm_fwd(B, d.ni, d.ni, d.g2, d.nip, nullptr, d.nip, P + d.co, d.xh, d.nip, ACT_SIGMOID);
  dec.push_back(make_gemm(c, "fwd_out", m, KIND_NN, f_o, d.Vo, false));

  // ---------------- decoder backward ----------------------------------------------------------------------------------
  bd.push_back(make_gemm(c, "wgrad_out", m, KIND_TN, gemm_wgrad(B, d.ni, d.ni, d.g2, d.nip, d.da, d.nip, G + d.Vo, d.nip, G + d.co), -1, false));
  bd.push_back(make_gemm(c, "dgrad_out", m, KIND_NT, gemm_dgrad(B, d.ni, d.ni, d.da, d.nip, nullptr, d.nip, d.dg2, d.nip, d.g2, d.nip, ACT_SIGMOID), d.Vo, R));
  // deconv4: d4 = dg2 [B,28,28,1]
  bd.push_back(make_colsum(c, "bgrad_deconv4", m, d.dg2, 1, (int64_t)B * d.ni, 1, G + d.d4));
  bd.push_back(make_im2col("bwd_im2col4", m, d.dg2, B, d.t4, 1, 5, 2, pb_s, d.t3, d.cols4, 28));
  bd.push_back(make_gemm(c, "wgrad_deconv4", m, KIND_TN, gemm_wgrad(m3, 25, d.q3, d.cols4, 28, d.o3, d.q3, G + d.D4, d.q3, nullptr), -1, false));
  { GemmArgs a = gemm_plain(m3, d.q3, 25, d.cols4, 28, d.q3, d.dd3, d.q3); a.aux = d.o3; a.ldaux = d.q3; a.act = ACT_SIGMOID;
    bd.push_back(make_gemm(c, "dgrad_deconv4", m, KIND_NN, a, d.D4, R)); }
  bd.push_back(make_colsum(c, "bgrad_deconv3", m, d.dd3, d.q3, m3, d.q3, G + d.d3));
  bd.push_back(make_im2col("bwd_im2col3", m, d.dd3, B, d.t3, d.q3, 5, 2, pb_s, d.t2, d.cols3, 25 * d.q3));
  bd.push_back(make_gemm(c, "wgrad_deconv3", m, KIND_TN, gemm_wgrad(m2, 25 * d.q3, d.q2, d.cols3, 25 * d.q3, d.o2, d.q2, G + d.D3, d.q2, nullptr), -1, false));
  { GemmArgs a = gemm_plain(m2, d.q2, 25 * d.q3, d.cols3, 25 * d.q3, d.q2, d.dd2, d.q2); a.aux = d.o2; a.ldaux = d.q2; a.act = ACT_SIGMOID;
    bd.push_back(make_gemm(c, "dgrad_deconv3", m, KIND_NN, a, d.D3, R)); }
  bd.push_back(make_colsum(c, "bgrad_deconv2", m, d.dd2, d.q2, m2, d.q2, G + d.d2));
  bd.push_back(make_im2col("bwd_im2col2", m, d.dd2, B, d.t2, d.q2, 5, 1, 0, d.t1, d.cols2, 25 * d.q2));
  bd.push_back(make_gemm(c, "wgrad_deconv2", m, KIND_TN, gemm_wgrad(m1, 25 * d.q2, d.q1, d.cols2, 25 * d.q2, d.o1, d.q1, G + d.D2, d.q1, nullptr), -1, false));
  { GemmArgs a = gemm_plain(m1, d.q1, 25 * d.q2, d.cols2, 25 * d.q2, d.q1, d.dd1, d.q1); a.aux = d.o1; a.ldaux = d.q1; a.act = ACT_SIGMOID;
    bd.push_back(make_gemm(c, "dgrad_deconv2", m, KIND_NN, a, d.D2, R)); }
  bd.push_back(make_colsum(c, "bgrad_deconv1", m, d.dd1, d.q1, m1, d.q1, G + d.d1));
  // deconv1 maps the 1x1 latent "image" to 3x3: its patch matrix is dd1 itself, viewed [B, 9 g1]
  bd.push_back(make_gemm(c, "wgrad_deconv1", m, KIND_TN, gemm_wgrad(B, 9 * d.q1, nz, d.dd1, 9 * d.q1, d.z, nz, G + d.D1, nzp, nullptr), -1, false));
  bd.push_back(make_gemm(c, "dgrad_deconv1", m, KIND_NN, gemm_plain(B, nz, 9 * d.q1, d.dd1, 9 * d.q1, nzp, d.dz, nz), d.D1, false));

  // ---------------- encoder backward (linear convs: no activation gradient, no bias) ---------------------------------
  be.push_back(make_gemm(c, "wgrad_heads", m, KIND_TN, gemm_wgrad(B, d.r2, d.nh, d.h2, d.r2p, d.dhd, d.nh, G + d.Wh, d.nhp, G + d.bh), -1, false));
  be.push_back(make_gemm(c, "dgrad_heads", m, KIND_NT, gemm_dgrad(B, d.r2, d.nh, d.dhd, d.nh, nullptr, d.nhp, d.dh2, d.r2p, nullptr, 0, ACT_NONE), d.Wh, R));
  be.push_back(make_gemm(c, "wgrad_conv3", m, KIND_TN, gemm_wgrad(n3, 25 * d.e2, d.e3, d.P3, 25 * d.e2, d.dh2, d.e3, G + d.C3, d.e3, nullptr), -1, false));
  be.push_back(make_gemm(c, "dcols_conv3", m, KIND_NT, gemm_plain(n3, 25 * d.e2, d.e3, d.dh2, d.e3, d.e3, d.cols2, 25 * d.e2), d.C3, false));
  be.push_back(make_col2im("dgrad_conv3", m, d.cols2, 25 * d.e2, B, d.s2, d.e2, d.s3, 5, 1, 0, nullptr, ACT_NONE, R, d.dl1));
  be.push_back(make_gemm(c, "wgrad_conv2", m, KIND_TN, gemm_wgrad(n2, 25 * d.e1, d.e2, d.P2, 25 * d.e1, d.dl1, d.e2, G + d.C2, d.e2, nullptr), -1, false));
  be.push_back(make_gemm(c, "dcols_conv2", m, KIND_NT, gemm_plain(n2, 25 * d.e1, d.e2, d.dl1, d.e2, d.e2, d.cols3, 25 * d.e1), d.C2, false));
  be.push_back(make_col2im("dgrad_conv2", m, d.cols3, 25 * d.e1, B, d.s1, d.e1, d.s2, 5, 2, same_pad_before(d.s1, 5, 2), nullptr, ACT_NONE, R, d.dl05));
  be.push_back(make_gemm(c, "wgrad_conv1", m, KIND_TN, gemm_wgrad(n1, 25, d.e1, d.P1, 28, d.dl05, d.e1, G + d.C1, d.e1, nullptr), -1, false));

  for (auto& op : enc) c->ops_fwd_enc.push_back(op);
  for (auto& op : dec) c->ops_fwd_dec.push_back(op);
  for (auto& op : bd) c->ops_bwd_dec.push_back(op);
  for (auto& op : be) c->ops_bwd_enc.push_back(op);
}

void build_segments(Ctx* c);

void build_ops(Ctx* c) {
  destroy_graphs(c);
  c->free_op_ws();
  group_destroy(c->gplan);
  c->gplan = group_create();
  // z has pitch n_z, the heads' gradient 2 n_z: 64-deep k-blocks need both to be <= 32 or multiples of 32 floats
  group_set_deep_k(c->gplan, c->cfg.n_z <= 16 || c->cfg.n_z % 32 == 0);
  c->fused = false;
  c->ops_colsum_dec.clear(); c->ops_colsum_enc.clear();
  c->ops_fwd_enc.clear(); c->ops_latent_fwd.clear(); c->ops_fwd_dec.clear(); c->ops_loss.clear();
  c->ops_bwd_dec.clear(); c->ops_latent_bwd.clear(); c->ops_bwd_enc.clear();
  const int M = c->cfg.n_modalities;
  c->ops_enc_mod.assign(M, {}); c->ops_dec_mod.assign(M, {});
  c->ops_loss_mod.assign(M, {}); c->ops_bwd_dec_mod.assign(M, {}); c->ops_bwd_enc_mod.assign(M, {});
  const int B = c->cfg.batch_size, nz = c->cfg.n_z;
  const int f = fwd_act(c);
  const bool tf32 = c->cfg.precision == VAEASSOC_TF32;
  const float inv_bg = 1.0f / (float)global_batch(c);
  float* P = c->p;   // biases are always read from the fp32 master
  bool round_z = false, round_dheads = false, round_x = false;

  auto add_recon = [&](int m, bool round_da) {
    Mod& d = c->mods[m];
    Op op; op.name = "recon_loss." + std::to_string(m);
    ReconArgs a;
    a.batch = B; a.n_input = d.ni; a.binary = d.cfg.binary; a.slot = 2 * m;
    a.scale = d.cfg.binary ? d.cfg.weight * inv_bg : d.cfg.weight;
    a.x = d.xs; a.ldx = d.nip; a.xhat = d.xh; a.ldxh = d.nip; a.da = d.da; a.ldda = d.nip;
    a.row_loss = d.rec_loss; a.partials = d.partials;
    a.round_tf32 = round_da ? 1 : 0;
    c->recon_round_da[m] = round_da;
    d.recon_blocks = (int)std::min<int64_t>(std::max<int64_t>((B + 7) / 8, 1), kMaxPartialBlocks);
    op.bytes = 4.0 * 3 * B * d.ni;
    op.run = [a](cudaStream_t s) { launch_recon_loss(a, s); };
    c->ops_loss.push_back(op);
    c->ops_loss_mod[m].push_back(op);
  };

  for (int m = 0; m < M; ++m) {
    Mod& d = c->mods[m];
    c->masks_in_use[m] = false;
    if (d.conv) {
      build_ops_conv(c, m);
      add_recon(m, tf32);
      round_x = round_x || tf32;
      continue;
    }
    // argument sets of every dense contraction of this modality
    GemmArgs f_e1 = gemm_fwd(B, d.r1, d.ni, d.xs, d.nip, nullptr, d.r1p, P + d.b1, d.h1, d.r1p, f);
    GemmArgs f_e2 = gemm_fwd(B, d.r2, d.r1, d.h1, d.r1p, nullptr, d.r2p, P + d.b2, d.h2, d.r2p, f);
    GemmArgs f_hd = gemm_fwd(B, d.nh, d.r2, d.h2, d.r2p, nullptr, d.nhp, P + d.bh, d.hd, d.nh, ACT_NONE);
    GemmArgs f_d1 = gemm_fwd(B, d.r1, nz, d.z, nz, nullptr, d.r1p, P + d.c1, d.g1, d.r1p, f);
    GemmArgs f_d2 = gemm_fwd(B, d.r2, d.r1, d.g1, d.r1p, nullptr, d.r2p, P + d.c2, d.g2, d.r2p, f);
    GemmArgs f_o = gemm_fwd(B, d.ni, d.r2, d.g2, d.r2p, nullptr, d.nip, P + d.co, d.xh, d.nip, d.cfg.binary ? ACT_SIGMOID : ACT_NONE);
    GemmArgs w_o = gemm_wgrad(B, d.r2, d.ni, d.g2, d.r2p, d.da, d.nip, c->g + d.Vo, d.nip, c->g + d.co);
    GemmArgs d_o = gemm_dgrad(B, d.r2, d.ni, d.da, d.nip, nullptr, d.nip, d.dg2, d.r2p, d.g2, d.r2p, f);
    GemmArgs w_d2 = gemm_wgrad(B, d.r1, d.r2, d.g1, d.r1p, d.dg2, d.r2p, c->g + d.V2, d.r2p, c->g + d.c2);
    GemmArgs d_d2 = gemm_dgrad(B, d.r1, d.r2, d.dg2, d.r2p, nullptr, d.r2p, d.dg1, d.r1p, d.g1, d.r1p, f);
    GemmArgs w_d1 = gemm_wgrad(B, nz, d.r1, d.z, nz, d.dg1, d.r1p, c->g + d.V1, d.r1p, c->g + d.c1);
    GemmArgs d_d1 = gemm_dgrad(B, nz, d.r1, d.dg1, d.r1p, nullptr, d.r1p, d.dz, nz, nullptr, 0, ACT_NONE);
    GemmArgs w_hd = gemm_wgrad(B, d.r2, d.nh, d.h2, d.r2p, d.dhd, d.nh, c->g + d.Wh, d.nhp, c->g + d.bh);
    GemmArgs d_hd = gemm_dgrad(B, d.r2, d.nh, d.dhd, d.nh, nullptr, d.nhp, d.dh2, d.r2p, d.h2, d.r2p, f);
    GemmArgs w_e2 = gemm_wgrad(B, d.r1, d.r2, d.h1, d.r1p, d.dh2, d.r2p, c->g + d.W2, d.r2p, c->g + d.b2);
    GemmArgs d_e2 = gemm_dgrad(B, d.r1, d.r2, d.dh2, d.r2p, nullptr, d.r2p, d.dh1, d.r1p, d.h1, d.r1p, f);
    GemmArgs w_e1 = gemm_wgrad(B, d.ni, d.r1, d.xs, d.nip, d.dh1, d.r1p, c->g + d.W1, d.r1p, c->g + d.b1);
    // an output is rounded to tf32 by its producer iff one of its consumers runs on the tensor cores
    auto tc = [&](int kind, GemmArgs a, int64_t w_off) {
      if (w_off >= 0) a.B = c->p_tf32 + w_off;
      return tf32 && tc_supported(kind, a);
    };
    if (f == ACT_RELU && tc(KIND_NN, f_e1, d.W1) && tc(KIND_NN, f_e2, d.W2) && tc(KIND_NN, f_d1, d.V1) && tc(KIND_NN, f_d2, d.V2) &&
        tc(KIND_NT, d_o, d.Vo) && tc(KIND_NT, d_d2, d.V2) && tc(KIND_NT, d_hd, d.Wh) && tc(KIND_NT, d_e2, d.W2) &&
        !getenv("VAEASSOC_NO_MASK")) {
      // relu' needs one bit per element: producers (forward) and consumers (dgrad) are all tcgen05 tasks
      const int64_t w1 = (d.r1 + 31) / 32, w2 = (d.r2 + 31) / 32;
      f_e1.mask_out = d.mh1; f_e1.ldmask = w1;  d_e2.mask_in = d.mh1; d_e2.ldmask = w1;
      f_e2.mask_out = d.mh2; f_e2.ldmask = w2;  d_hd.mask_in = d.mh2; d_hd.ldmask = w2;
      f_d1.mask_out = d.mg1; f_d1.ldmask = w1;  d_d2.mask_in = d.mg1; d_d2.ldmask = w1;
      f_d2.mask_out = d.mg2; f_d2.ldmask = w2;  d_o.mask_in = d.mg2;  d_o.ldmask = w2;
      c->masks_in_use[m] = true;
    }
    const bool r_h1 = tc(KIND_NN, f_e2, d.W2) || tc(KIND_TN, w_e2, -1);
    const bool r_h2 = tc(KIND_NN, f_hd, d.Wh) || tc(KIND_TN, w_hd, -1);
    const bool r_g1 = tc(KIND_NN, f_d2, d.V2) || tc(KIND_TN, w_d2, -1);
    const bool r_g2 = tc(KIND_NN, f_o, d.Vo) || tc(KIND_TN, w_o, -1);
    const bool r_dg2 = tc(KIND_TN, w_d2, -1) || tc(KIND_NT, d_d2, d.V2);
    const bool r_dg1 = tc(KIND_TN, w_d1, -1) || tc(KIND_NT, d_d1, d.V1);
    const bool r_dh2 = tc(KIND_TN, w_e2, -1) || tc(KIND_NT, d_e2, d.W2);
    const bool r_dh1 = tc(KIND_TN, w_e1, -1);
    round_z = round_z || tc(KIND_NN, f_d1, d.V1) || tc(KIND_TN, w_d1, -1);
    round_x = round_x || tc(KIND_NN, f_e1, d.W1) || tc(KIND_TN, w_e1, -1);
    round_dheads = round_dheads || tc(KIND_TN, w_hd, -1) || tc(KIND_NT, d_hd, d.Wh);

    auto& enc = c->ops_enc_mod[m];
    enc.push_back(make_gemm(c, "fwd_enc1", m, KIND_NN, f_e1, d.W1, r_h1));
    enc.push_back(make_gemm(c, "fwd_enc2", m, KIND_NN, f_e2, d.W2, r_h2));
    enc.push_back(make_gemm(c, "fwd_heads", m, KIND_NN, f_hd, d.Wh, false));
    auto& dec = c->ops_dec_mod[m];
    dec.push_back(make_gemm(c, "fwd_dec1", m, KIND_NN, f_d1, d.V1, r_g1));
    dec.push_back(make_gemm(c, "fwd_dec2", m, KIND_NN, f_d2, d.V2, r_g2));
    dec.push_back(make_gemm(c, "fwd_out", m, KIND_NN, f_o, d.Vo, false));
    for (auto& op : enc) c->ops_fwd_enc.push_back(op);
    for (auto& op : dec) c->ops_fwd_dec.push_back(op);

    auto& bd = c->ops_bwd_dec_mod[m];
    bd.push_back(make_gemm(c, "wgrad_out", m, KIND_TN, w_o, -1, false));
    bd.push_back(make_gemm(c, "dgrad_out", m, KIND_NT, d_o, d.Vo, r_dg2));
    bd.push_back(make_gemm(c, "wgrad_dec2", m, KIND_TN, w_d2, -1, false));
    bd.push_back(make_gemm(c, "dgrad_dec2", m, KIND_NT, d_d2, d.V2, r_dg1));
    bd.push_back(make_gemm(c, "wgrad_dec1", m, KIND_TN, w_d1, -1, false));
    bd.push_back(make_gemm(c, "dgrad_dec1", m, KIND_NT, d_d1, d.V1, false));
    auto& be = c->ops_bwd_enc_mod[m];
    be.push_back(make_gemm(c, "wgrad_heads", m, KIND_TN, w_hd, -1, false));
    be.push_back(make_gemm(c, "dgrad_heads", m, KIND_NT, d_hd, d.Wh, r_dh2));
    be.push_back(make_gemm(c, "wgrad_enc2", m, KIND_TN, w_e2, -1, false));
    be.push_back(make_gemm(c, "dgrad_enc2", m, KIND_NT, d_e2, d.W2, r_dh1));
    be.push_back(make_gemm(c, "wgrad_enc1", m, KIND_TN, w_e1, -1, false));
    for (auto& o : bd) c->ops_bwd_dec.push_back(o);
    for (auto& o : be) c->ops_bwd_enc.push_back(o);

    add_recon(m, tc(KIND_TN, w_o, -1) || tc(KIND_NT, d_o, d.Vo));
  }
  {
    Op op; op.name = "latent_fwd";
    LatentArgs a;
    a.n_mod = M; a.batch = B; a.n_z = nz; a.inv_global_batch = inv_bg; a.lambda = c->cfg.assoc_lambda;
    for (int m = 0; m < M; ++m) {
      Mod& d = c->mods[m];
      a.weight[m] = d.cfg.weight; a.heads[m] = d.hd; a.z[m] = d.z; a.gstat[m] = d.gstat; a.latent_loss[m] = d.lat_loss;
    }
    a.eps = c->eps; a.partials = c->lat_partials; a.with_grad = 1; a.round_z = round_z ? 1 : 0;
    c->lat_blocks = (int)std::min<int64_t>(std::max<int64_t>((B + 255) / 256, 1), kMaxPartialBlocks);
    c->round_z = round_z; c->round_x = round_x;
    op.bytes = 4.0 * B * nz * (1 + M * 5);
    op.run = [a](cudaStream_t s) { launch_latent_fwd(a, s); };
    c->ops_latent_fwd.push_back(op);
    c->lat_fwd_args = a;
  }
  {
    Op op; op.name = "latent_bwd";
    LatentBwdArgs a;
    a.n_mod = M; a.batch = B; a.n_z = nz; a.eps = c->eps; a.round_out = round_dheads ? 1 : 0;
    for (int m = 0; m < M; ++m) {
      Mod& d = c->mods[m];
      a.heads[m] = d.hd; a.gstat[m] = d.gstat; a.dz[m] = d.dz; a.dheads[m] = d.dhd;
    }
    op.bytes = 4.0 * B * nz * (1 + M * 6);
    op.run = [a](cudaStream_t s) { launch_latent_bwd(a, s); };
    c->ops_latent_bwd.push_back(op);
    c->lat_bwd_args = a;
  }
  build_segments(c);
}

// ---- fused train-step segments (dense modalities, tf32) -------------------------------------------------------------
// The layers of a segment become ONE launch of the persistent tile kernel: tasks in dependency order, linked by
// row-block counters (tensor T of modality m over rows [256 rb, +256) is complete when its counter reaches
// 16 x (N-tiles of the producing layer)).  Counter index = (T * n_modalities + m) * RB + rb.
// (forward tensors first, then backward ones: a launch rewinds one contiguous range of counters when it leaves)
enum { T_H1 = 0, T_H2, T_HD, T_Z, T_G1, T_G2, T_DG2, T_DG1, T_DZ, T_DHD, T_DH2, T_DH1, T_DA, T_COUNT };
static_assert(T_COUNT == 13, "alloc_buffers sizes the counter array for 13 tensors per modality");

void build_segments(Ctx* c) {
  const int M = c->cfg.n_modalities;
  const int B = c->cfg.batch_size;
  const int RB = (B + 255) / 256;
  bool ok = c->cfg.precision == VAEASSOC_TF32 && !getenv("VAEASSOC_NO_FUSE");
  const bool nodeps = getenv("VAEASSOC_DEBUG_NODEPS") != nullptr;   // timing experiments only: wrong results
  const bool half_ok = kGroupHalfOk && getenv("VAEASSOC_NO_HALF") == nullptr;       // half-tile hand-over between dependent layers (one-launch form)
  for (int m = 0; m < M && ok; ++m) {
    if (c->mods[m].conv) { ok = false; break; }
    if (c->ops_enc_mod[m].size() != 3 || c->ops_dec_mod[m].size() != 3 || c->ops_bwd_dec_mod[m].size() != 6 ||
        c->ops_bwd_enc_mod[m].size() != 5) { ok = false; break; }
    for (auto* v : {&c->ops_enc_mod[m], &c->ops_dec_mod[m], &c->ops_bwd_dec_mod[m], &c->ops_bwd_enc_mod[m]})
      for (const Op& op : *v) ok = ok && op.kind >= 0;
  }
  char err[256] = {0};
  if (ok) {
    GroupPlan* g = c->gplan;
    auto ctr = [&](int m, int T, int rb) { return (T * M + m) * RB + rb; };
    // tiles of a row-wise layer (NN / NT): one task per (row block, column tile)
    // in_w > 0: the input tensor comes from row-wise tiles in_w k-blocks wide that publish their first half early
    // (half-tile hand-over, gemm_group.cu TF_HALF); sig_half: this layer's tiles do so for their consumer; *out_w = this
    // layer's tile width in k-blocks of its consumer when the hand-over pays (tiles of >= 4 chunks), else 0
    auto add_rowwise = [&](const Op& op, int m, int inT, int in_tn, int outT, float* colsum, int in_m = -1,
                           const GemmArgs* override_args = nullptr, int in_w = 0, bool sig_half = false,
                           int* out_w = nullptr) -> int {
      if (nodeps) { inT = -1; outT = -1; }
      if (in_m < 0) in_m = m;
      GemmArgs a = override_args ? *override_args : op.gargs;
      a.bias_grad = getenv("VAEASSOC_DEBUG_NO_COLSUM") ? nullptr : colsum;     // timing experiments only: no bias gradients
      const int prob = group_add_problem(g, op.kind, a, err, sizeof err);
      if (prob < 0) fail("segment plan for %s failed: %s", op.name.c_str(), err);
      const int tm = group_problem_tiles_m(g, prob), tn = group_problem_tiles_n(g, prob), kb = group_problem_kblocks(g, prob);
      const int bn = group_problem_bn(g, prob);
      const bool sig = sig_half && outT >= 0 && bn >= 128 && half_ok;
      const bool half = in_w > 0 && inT >= 0 && half_ok;
      if (out_w) *out_w = sig ? bn / 32 : 0;
      for (int i = 0; i < tm; ++i)
        for (int j = 0; j < tn; ++j)
          group_add_task(g, prob, i, j, 0, kb, inT >= 0 ? ctr(in_m, inT, i) : -1, inT >= 0 ? 1 : 0,
                         kGroupSignalsPerTile * in_tn, half ? ctr(in_m, inT, i) + c->n_ctr_half : -1, half ? in_w : 0,
                         outT >= 0 ? ctr(m, outT, i) : -1, (half ? kTaskHalf : 0) | (sig ? kTaskSigHalf : 0));
      return tn;
    };
    // weight gradient (TN): the batch contraction is cut into row-block ranges; a task waits for dY over its range
    auto add_wgrad = [&](const Op& op, int m, int dyT, int dy_tn, int dy_m = -1, bool want_colsum_op = true) {
      const bool external = dyT < 0 && want_colsum_op;
      if (nodeps) dyT = -1;
      if (dy_m < 0) dy_m = m;
      GemmArgs a = op.gargs;
      a.bias_grad = nullptr;
      const int prob = group_add_problem(g, op.kind, a, err, sizeof err);
      if (prob < 0) fail("segment plan for %s failed: %s", op.name.c_str(), err);
      const int tm = group_problem_tiles_m(g, prob), tn = group_problem_tiles_n(g, prob), kb = group_problem_kblocks(g, prob);
      const int kpr = group_problem_kb_per_rowblock(g, prob);
      const int rbs = (kb + kpr - 1) / kpr;
      const int splits = std::max(1, std::min(rbs, (kNumSMs / 2) / std::max(1, tm * tn)));
      const int per = (rbs + splits - 1) / splits;
      for (int r0 = 0; r0 < rbs; r0 += per) {
        const int nrb = std::min(per, rbs - r0);
        for (int i = 0; i < tm; ++i)
          for (int j = 0; j < tn; ++j)
            group_add_task(g, prob, i, j, r0 * kpr, std::min(nrb * kpr, kb - r0 * kpr), dyT >= 0 ? ctr(dy_m, dyT, r0) : -1,
                           dyT >= 0 ? nrb : 0, kGroupSignalsPerTile * dy_tn, -1, 0, -1);
      }
      if (external && op.gargs.bias_grad) {
        // dY comes from an elementwise kernel (loss / latent backward): its column sums need their own launch
        const GemmArgs b = op.gargs;
        Op cs; cs.name = "bgrad_" + op.name; cs.bytes = 4.0 * b.K * b.N;
        float* cws = c->op_ws(colsum_ws_floats(b.K, b.N));
        cs.run = [b, cws](cudaStream_t s) { launch_colsum(b.B, b.ldb, b.K, b.N, b.bias_grad, cws, s); };
        return cs;
      }
      return Op();
    };
    // a row-wise layer with a tiny N and a long K (heads: N = 2 n_z; decoder input dgrad: N = n_z) as split-K tasks: every
    // k range is its own tile task (raw accumulator, TMA reduce-add into C, which the staging kernel cleared; bias of
    // the heads: added by the latent task).  The layer's stage costs a quarter of its main loop instead of all of it.
    // Returns the number of tasks per row block (the consumer waits for kGroupSignalsPerTile x that).
    auto add_rowwise_splitk = [&](const Op& op, int m, int inT, int in_tn, int outT, int in_w) -> int {
      GemmArgs a = op.gargs;
      a.bias = nullptr; a.bias_grad = nullptr; a.force_reduce = 1; a.act = ACT_NONE; a.round_out = 0;
      a.aux = nullptr; a.mask_in = nullptr; a.mask_out = nullptr;
      const int prob = group_add_problem(g, op.kind, a, err, sizeof err);
      if (prob < 0) fail("segment plan for %s failed: %s", op.name.c_str(), err);
      const int tm = group_problem_tiles_m(g, prob), tn = group_problem_tiles_n(g, prob), kb = group_problem_kblocks(g, prob);
      const int splits = std::max(1, std::min(4, kb / 3));
      const int per = (kb + splits - 1) / splits;
      const bool half = in_w > 0 && inT >= 0 && half_ok && splits == 1;
      int n_tasks = 0;
      for (int i = 0; i < tm; ++i) {
        n_tasks = 0;
        for (int k0 = 0; k0 < kb; k0 += per)
          for (int j = 0; j < tn; ++j, ++n_tasks)
            group_add_task(g, prob, i, j, k0, std::min(per, kb - k0), inT >= 0 ? ctr(m, inT, i) : -1, inT >= 0 ? 1 : 0,
                           kGroupSignalsPerTile * in_tn, half ? ctr(m, inT, i) + c->n_ctr_half : -1, half ? in_w : 0,
                           outT >= 0 ? ctr(m, outT, i) : -1, half ? kTaskHalf : 0);
      }
      return n_tasks;
    };
    auto begin = [&](Ctx::Seg& sg, int T0, int n_tensors = 2) {
      sg.site = group_begin(g);
      if (sg.site >= c->max_sites - 1) fail("too many tensor-core launch sites");
      sg.reset_first = ctr(0, T0, 0);
      sg.reset_count = n_tensors * M * RB;
    };
    auto end = [&](Ctx::Seg& sg) { (void)sg; if (!group_end(g, err, sizeof err)) fail("%s", err); };
    std::vector<int> tn1(M), tn2(M);
    // encoder forward: enc1 -> h1 -> enc2 -> h2 -> heads
    begin(c->seg_enc, T_H1);
    for (int m = 0; m < M; ++m) tn1[m] = add_rowwise(c->ops_enc_mod[m][0], m, -1, 0, T_H1, nullptr);
    for (int m = 0; m < M; ++m) tn2[m] = add_rowwise(c->ops_enc_mod[m][1], m, T_H1, tn1[m], T_H2, nullptr);
    for (int m = 0; m < M; ++m) add_rowwise(c->ops_enc_mod[m][2], m, T_H2, tn2[m], -1, nullptr);
    end(c->seg_enc);
    // decoder forward: dec1 -> g1 -> dec2 -> g2 -> out
    begin(c->seg_dec, T_G1);
    for (int m = 0; m < M; ++m) tn1[m] = add_rowwise(c->ops_dec_mod[m][0], m, -1, 0, T_G1, nullptr);
    for (int m = 0; m < M; ++m) tn2[m] = add_rowwise(c->ops_dec_mod[m][1], m, T_G1, tn1[m], T_G2, nullptr);
    for (int m = 0; m < M; ++m) add_rowwise(c->ops_dec_mod[m][2], m, T_G2, tn2[m], -1, nullptr);
    end(c->seg_dec);
    // decoder backward; ops: 0 wgrad_out 1 dgrad_out 2 wgrad_dec2 3 dgrad_dec2 4 wgrad_dec1 5 dgrad_dec1.  The dgrad
    // epilogue that produces dY also produces the bias gradient (column sums) of the layer that consumes dY
    begin(c->seg_bwd_dec, T_DG2);
    for (int m = 0; m < M; ++m) { auto& bd = c->ops_bwd_dec_mod[m]; tn1[m] = add_rowwise(bd[1], m, -1, 0, T_DG2, bd[2].gargs.bias_grad); }
    for (int m = 0; m < M; ++m) { Op cs = add_wgrad(c->ops_bwd_dec_mod[m][0], m, -1, 0); if (cs.run) c->ops_colsum_dec.push_back(cs); }
    for (int m = 0; m < M; ++m) { auto& bd = c->ops_bwd_dec_mod[m]; tn2[m] = add_rowwise(bd[3], m, T_DG2, tn1[m], T_DG1, bd[4].gargs.bias_grad); }
    for (int m = 0; m < M; ++m) add_wgrad(c->ops_bwd_dec_mod[m][2], m, T_DG2, tn1[m]);
    for (int m = 0; m < M; ++m) add_rowwise(c->ops_bwd_dec_mod[m][5], m, T_DG1, tn2[m], -1, nullptr);
    for (int m = 0; m < M; ++m) add_wgrad(c->ops_bwd_dec_mod[m][4], m, T_DG1, tn2[m]);
    end(c->seg_bwd_dec);
    // encoder backward; ops: 0 wgrad_heads 1 dgrad_heads 2 wgrad_enc2 3 dgrad_enc2 4 wgrad_enc1
    begin(c->seg_bwd_enc, T_DH2);
    for (int m = 0; m < M; ++m) { auto& be = c->ops_bwd_enc_mod[m]; tn1[m] = add_rowwise(be[1], m, -1, 0, T_DH2, be[2].gargs.bias_grad); }
    for (int m = 0; m < M; ++m) { Op cs = add_wgrad(c->ops_bwd_enc_mod[m][0], m, -1, 0); if (cs.run) c->ops_colsum_enc.push_back(cs); }
    for (int m = 0; m < M; ++m) { auto& be = c->ops_bwd_enc_mod[m]; tn2[m] = add_rowwise(be[3], m, T_DH2, tn1[m], T_DH1, be[4].gargs.bias_grad); }
    for (int m = 0; m < M; ++m) add_wgrad(c->ops_bwd_enc_mod[m][2], m, T_DH2, tn1[m]);
    for (int m = 0; m < M; ++m) add_wgrad(c->ops_bwd_enc_mod[m][4], m, T_DH1, tn2[m]);
    end(c->seg_bwd_enc);
    c->fused = true;
    // ---- two-launch form: latent stages as elementwise tasks between the encoder and decoder layers ----
    c->elt_built = false;
    // (a latent task is one thread per batch row, serial over n_z: at n_z = 64 it costs more on the critical path of its
    // row block than the stand-alone kernel spread over all SMs -- measured 3.86 against 3.52 ms per step at the scaled
    // config -- so wide latents keep the four-launch form)
    if (M <= 2 && c->cfg.n_z <= 16 && RB * 8 <= kMaxPartialBlocks && !nodeps && !getenv("VAEASSOC_NO_ELT")) {
      GElem el;
      el.lf = c->lat_fwd_args;
      el.lb = c->lat_bwd_args;
      for (int m = 0; m < M; ++m) el.bh_grad[m] = c->ops_bwd_enc_mod[m][0].gargs.bias_grad;
      group_set_elem(g, el);
      c->lat_blocks_elt = RB * 8;
      const int S = kGroupSignalsPerTile;
      std::vector<int> tnh(M), tnz(M);
      begin(c->seg_fwd, T_H1, 6);
      for (int m = 0; m < M; ++m) tn1[m] = add_rowwise(c->ops_enc_mod[m][0], m, -1, 0, T_H1, nullptr);
      for (int m = 0; m < M; ++m) tn2[m] = add_rowwise(c->ops_enc_mod[m][1], m, T_H1, tn1[m], T_H2, nullptr);
      for (int m = 0; m < M; ++m) tnh[m] = add_rowwise(c->ops_enc_mod[m][2], m, T_H2, tn2[m], T_HD, nullptr);
      for (int rb = 0; rb < RB; ++rb)
        group_add_elt_task(g, 0, rb, B, ctr(0, T_HD, rb), 1, S * tnh[0], M > 1 ? ctr(1, T_HD, rb) : -1, M > 1 ? S * tnh[1] : 0,
                           ctr(0, T_Z, rb));
      for (int m = 0; m < M; ++m) tn1[m] = add_rowwise(c->ops_dec_mod[m][0], m, T_Z, 1, T_G1, nullptr, 0);
      for (int m = 0; m < M; ++m) tn2[m] = add_rowwise(c->ops_dec_mod[m][1], m, T_G1, tn1[m], T_G2, nullptr);
      for (int m = 0; m < M; ++m) add_rowwise(c->ops_dec_mod[m][2], m, T_G2, tn2[m], -1, nullptr);
      end(c->seg_fwd);
      begin(c->seg_bwd, T_DG2, 6);
      for (int m = 0; m < M; ++m) { auto& bd = c->ops_bwd_dec_mod[m]; tn1[m] = add_rowwise(bd[1], m, -1, 0, T_DG2, bd[2].gargs.bias_grad); }
      for (int m = 0; m < M; ++m) add_wgrad(c->ops_bwd_dec_mod[m][0], m, -1, 0);     // (its column sums: ops_colsum_dec, built above)
      for (int m = 0; m < M; ++m) { auto& bd = c->ops_bwd_dec_mod[m]; tn2[m] = add_rowwise(bd[3], m, T_DG2, tn1[m], T_DG1, bd[4].gargs.bias_grad); }
      for (int m = 0; m < M; ++m) add_wgrad(c->ops_bwd_dec_mod[m][2], m, T_DG2, tn1[m]);
      for (int m = 0; m < M; ++m) tnz[m] = add_rowwise(c->ops_bwd_dec_mod[m][5], m, T_DG1, tn2[m], T_DZ, nullptr);
      for (int m = 0; m < M; ++m) add_wgrad(c->ops_bwd_dec_mod[m][4], m, T_DG1, tn2[m]);
      for (int rb = 0; rb < RB; ++rb)
        group_add_elt_task(g, 1, rb, B, ctr(0, T_DZ, rb), 1, S * tnz[0], M > 1 ? ctr(1, T_DZ, rb) : -1, M > 1 ? S * tnz[1] : 0,
                           ctr(0, T_DHD, rb));
      for (int m = 0; m < M; ++m) { auto& be = c->ops_bwd_enc_mod[m]; tn1[m] = add_rowwise(be[1], m, T_DHD, 1, T_DH2, be[2].gargs.bias_grad, 0); }
      for (int m = 0; m < M; ++m) add_wgrad(c->ops_bwd_enc_mod[m][0], m, T_DHD, 1, 0);   // bias gradient: by the latent-backward task
      for (int m = 0; m < M; ++m) { auto& be = c->ops_bwd_enc_mod[m]; tn2[m] = add_rowwise(be[3], m, T_DH2, tn1[m], T_DH1, be[4].gargs.bias_grad); }
      for (int m = 0; m < M; ++m) add_wgrad(c->ops_bwd_enc_mod[m][2], m, T_DH2, tn1[m]);
      for (int m = 0; m < M; ++m) add_wgrad(c->ops_bwd_enc_mod[m][4], m, T_DH1, tn2[m]);
      end(c->seg_bwd);
      c->elt_built = true;
      // ---- one-launch form: the whole gradient step (forward, losses, cost, backward) is ONE launch ----
      // The output layer of each decoder turns its accumulator straight into the reconstruction loss and d cost / d a
      // (target tile through the epilogue's aux boxes), also summing the output bias gradient; the decoder backward
      // waits on those tiles (T_DA).  The cost reduction is one more elementwise task.
      c->one_built = false;
      if (!getenv("VAEASSOC_NO_ONE") && 17 * M <= 36) {
        std::vector<int> tno(M);
        // modalities in order of increasing width: within a layer the short tasks come first, so that the narrow modality's
        // chain (which the latent stage needs as much as the wide one's) never queues behind a wave of long main loops
        std::vector<int> mord(M);
        for (int m = 0; m < M; ++m) mord[m] = m;
        if (!getenv("VAEASSOC_ORDER_DECLARED"))
          std::stable_sort(mord.begin(), mord.end(), [&](int x, int y) { return c->mods[x].ni < c->mods[y].ni; });
        begin(c->seg_step, T_H1, T_COUNT);
        std::vector<int> w1(M), w2(M), wo(M);
        for (int m : mord) tn1[m] = add_rowwise(c->ops_enc_mod[m][0], m, -1, 0, T_H1, nullptr, -1, nullptr, 0, true, &w1[m]);
        for (int m : mord) tn2[m] = add_rowwise(c->ops_enc_mod[m][1], m, T_H1, tn1[m], T_H2, nullptr, -1, nullptr, w1[m], true, &w2[m]);
        // (split-K pays where a layer's main loop is its stage: one to three row blocks.  At 8192 pairs the extra tasks
        // cost more than the shorter stage gains: 0.283 against 0.276 ms per step; at 100 pairs 0.132 against 0.136)
        c->split_heads = RB < 4 && getenv("VAEASSOC_NO_SPLIT_HEADS") == nullptr;
        if (getenv("VAEASSOC_SPLIT_HEADS")) c->split_heads = true;
        for (int m : mord) {
          if (c->split_heads) { tnh[m] = add_rowwise_splitk(c->ops_enc_mod[m][2], m, T_H2, tn2[m], T_HD, w2[m]); el.lf.head_bias[m] = c->ops_enc_mod[m][2].gargs.bias; }
          else tnh[m] = add_rowwise(c->ops_enc_mod[m][2], m, T_H2, tn2[m], T_HD, nullptr, -1, nullptr, w2[m]);
        }
        for (int rb = 0; rb < RB; ++rb)
          group_add_elt_task(g, 0, rb, B, ctr(0, T_HD, rb), 1, S * tnh[0], M > 1 ? ctr(1, T_HD, rb) : -1, M > 1 ? S * tnh[1] : 0,
                             ctr(0, T_Z, rb), 1, c->split_heads ? 1 : 0);
        for (int m : mord) tn1[m] = add_rowwise(c->ops_dec_mod[m][0], m, T_Z, 1, T_G1, nullptr, 0, nullptr, 0, true, &w1[m]);
        for (int m : mord) tn2[m] = add_rowwise(c->ops_dec_mod[m][1], m, T_G1, tn1[m], T_G2, nullptr, -1, nullptr, w1[m], true, &w2[m]);
        FinalizeArgs fin = finalize_args(c, 0);
        fin.partials_latent = c->lat_partials; fin.blocks_latent = c->lat_blocks_elt;
        for (int m : mord) {
          Mod& d = c->mods[m];
          GemmArgs a = c->ops_dec_mod[m][2].gargs;           // x_hat = act(g2 Vo + co)  ->  d a, loss
          const GemmArgs& w_o = c->ops_bwd_dec_mod[m][0].gargs;
          const int tiles = (d.ni + 63) / 64 * RB;             // upper bound of the tile count (narrowest tile)
          float* parts = c->op_ws((int64_t)tiles * kGroupSignalsPerTile);
          a.C = d.da; a.ldc = d.nip; a.act = ACT_NONE;
          a.round_out = c->recon_round_da[m] ? 1 : 0;
          a.loss_x = d.xs; a.ld_loss_x = d.nip; a.loss_partials = parts;
          a.loss_scale = d.cfg.binary ? d.cfg.weight * (1.0f / (float)global_batch(c)) : d.cfg.weight;
          a.loss_binary = d.cfg.binary ? 1 : 0;
          tno[m] = add_rowwise(c->ops_dec_mod[m][2], m, T_G2, tn2[m], T_DA, w_o.bias_grad, -1, &a, w2[m], true, &wo[m]);
          fin.partials_recon[m] = parts; fin.blocks_recon[m] = RB * tno[m] * kGroupSignalsPerTile;
          fin.stride_recon[m] = 1; fin.off_recon[m] = 0;
        }
        c->fin_one = fin;
        el.fin = fin;
        group_set_elem(g, el);
        for (int m : mord) { auto& bd = c->ops_bwd_dec_mod[m]; tn1[m] = add_rowwise(bd[1], m, T_DA, tno[m], T_DG2, bd[2].gargs.bias_grad, -1, nullptr, wo[m], true, &w1[m]); }
        for (int m : mord) add_wgrad(c->ops_bwd_dec_mod[m][0], m, T_DA, tno[m]);
        for (int m : mord) { auto& bd = c->ops_bwd_dec_mod[m]; tn2[m] = add_rowwise(bd[3], m, T_DG2, tn1[m], T_DG1, bd[4].gargs.bias_grad, -1, nullptr, w1[m], true, &w2[m]); }
        group_add_elt_task(g, 2, 0, B, ctr(0, T_DA, 0), RB, S * tno[0], M > 1 ? ctr(1, T_DA, 0) : -1, M > 1 ? S * tno[1] : 0, -1, RB);
        for (int m : mord) add_wgrad(c->ops_bwd_dec_mod[m][2], m, T_DG2, tn1[m]);
        for (int m : mord) {
          if (c->split_heads) tnz[m] = add_rowwise_splitk(c->ops_bwd_dec_mod[m][5], m, T_DG1, tn2[m], T_DZ, w2[m]);
          else tnz[m] = add_rowwise(c->ops_bwd_dec_mod[m][5], m, T_DG1, tn2[m], T_DZ, nullptr, -1, nullptr, w2[m]);
        }
        for (int m : mord) add_wgrad(c->ops_bwd_dec_mod[m][4], m, T_DG1, tn2[m]);
        for (int rb = 0; rb < RB; ++rb)
          group_add_elt_task(g, 1, rb, B, ctr(0, T_DZ, rb), 1, S * tnz[0], M > 1 ? ctr(1, T_DZ, rb) : -1, M > 1 ? S * tnz[1] : 0,
                             ctr(0, T_DHD, rb));
        for (int m : mord) { auto& be = c->ops_bwd_enc_mod[m]; tn1[m] = add_rowwise(be[1], m, T_DHD, 1, T_DH2, be[2].gargs.bias_grad, 0, nullptr, 0, true, &w1[m]); }
        for (int m : mord) add_wgrad(c->ops_bwd_enc_mod[m][0], m, T_DHD, 1, 0);
        for (int m : mord) { auto& be = c->ops_bwd_enc_mod[m]; tn2[m] = add_rowwise(be[3], m, T_DH2, tn1[m], T_DH1, be[4].gargs.bias_grad, -1, nullptr, w1[m]); }
        for (int m : mord) add_wgrad(c->ops_bwd_enc_mod[m][2], m, T_DH2, tn1[m]);
        for (int m : mord) add_wgrad(c->ops_bwd_enc_mod[m][4], m, T_DH1, tn2[m]);
        end(c->seg_step);
        c->one_built = true;
      }
    }
  }
  group_set_counters(c->gplan, c->gsync, c->n_ctr, c->n_ctr_half);
  if (!group_upload(c->gplan, err, sizeof err)) fail("%s", err);
}

void launch_seg(Ctx* c, const Ctx::Seg& sg, cudaStream_t s, int advance = 0) {
  group_launch(c->gplan, sg.site, c->gsync + c->n_ctr + 2 * sg.site, sg.reset_first, sg.reset_count, dyn_first(c), s, advance);
  c->launches += 1;
}

void run_ops(Ctx* c, std::vector<Op>& ops, cudaStream_t s) {
  for (Op& op : ops) {
    op.run(s);
    c->launches += op.launches;
  }
}

// backward slice of modality m on stream s: the dgrad chain stays on s, every weight-gradient op forks onto the
// modality's wgrad stream at the point where its dY is complete; the branch joins s again at the end of the slice
// (under stream capture the event edges become graph dependencies)
void run_bwd_ops(Ctx* c, int m, std::vector<Op>& ops, cudaStream_t s) {
  cudaStream_t w = c->wstream[m];
  bool forked = false;
  for (Op& op : ops) {
    if (op.side) {
      CUDA_OK(cudaEventRecord(c->ev_wfork[m], s));
      CUDA_OK(cudaStreamWaitEvent(w, c->ev_wfork[m], 0));
      op.run(w);
      forked = true;
    } else {
      op.run(s);
    }
    c->launches += op.launches;
  }
  if (forked) {
    CUDA_OK(cudaEventRecord(c->ev_wjoin[m], w));
    CUDA_OK(cudaStreamWaitEvent(s, c->ev_wjoin[m], 0));
  }
}

// the two-launch form (latent stages inside the tile kernel) serves every schedule except the two-bucket NCCL one, whose
// first all-reduce starts between the decoder and the encoder backward
bool dp_two_buckets(const Ctx* c) {
  const bool dp = c->comm != nullptr && c->world > 1;
  if (!dp || c->peer.on) return false;
  return !(c->dp_single >= 0 ? c->dp_single != 0 : (c->fused && c->world > 4));
}
bool elt_mode(const Ctx* c) { return c->fused && c->elt_built && !dp_two_buckets(c); }
bool one_mode(const Ctx* c) { return elt_mode(c) && c->one_built; }

FinalizeArgs finalize_args(Ctx* c, int advance) {
  FinalizeArgs a;
  a.n_mod = c->cfg.n_modalities;
  for (int m = 0; m < a.n_mod; ++m) {
    a.binary[m] = c->mods[m].cfg.binary; a.weight[m] = c->mods[m].cfg.weight;
    a.partials_recon[m] = c->mods[m].partials; a.blocks_recon[m] = c->mods[m].recon_blocks;
  }
  a.inv_global_batch = 1.0f / (float)global_batch(c); a.lambda = c->cfg.assoc_lambda;
  a.partials_latent = c->lat_partials; a.blocks_latent = c->lat_mode_elt ? c->lat_blocks_elt : c->lat_blocks;
  a.scalars = c->scalars; a.cost_slot = c->g + c->n_flat; a.step_dev = c->step_dev; a.advance = advance;
  return a;
}

AdamArgs adam_args(Ctx* c) {
  AdamArgs a;
  a.p = c->p; a.g = c->g; a.m = c->m; a.v = c->v;
  a.p_tf32 = c->cfg.precision == VAEASSOC_TF32 ? c->p_tf32 : nullptr;
  a.n = c->n_flat; a.lr = c->cfg.learning_rate; a.beta1 = c->cfg.beta1; a.beta2 = c->cfg.beta2; a.eps = c->cfg.adam_epsilon;
  a.step_dev = c->step_dev; a.cost_slot = c->g + c->n_flat; a.cost_hist = c->cost_hist; a.hist_cap = c->hist_cap;
  a.last_cost = c->last_cost;
  return a;
}

// The modalities only meet in the latent kernels (shared eps, association KL): between those joins each modality's
// chain of layers runs on its own stream (modality 0 on `s`, modality m on side[m-1]).  Under stream capture the
// event edges become graph dependencies, so the replayed graph has one branch per modality; the small joint-modality
// GEMMs (10 % of the FLOPs, latency-bound) then hide inside the image-modality ones.
cudaStream_t mod_stream(Ctx* c, int m, cudaStream_t s) { return m == 0 ? s : c->side[m - 1]; }
void fork_modalities(Ctx* c, cudaStream_t s) {
  if (c->cfg.n_modalities < 2) return;
  CUDA_OK(cudaEventRecord(c->ev_fork, s));
  for (int m = 1; m < c->cfg.n_modalities; ++m) CUDA_OK(cudaStreamWaitEvent(c->side[m - 1], c->ev_fork, 0));
}
void join_modalities(Ctx* c, cudaStream_t s) {
  for (int m = 1; m < c->cfg.n_modalities; ++m) {
    CUDA_OK(cudaEventRecord(c->ev_join[m - 1], c->side[m - 1]));
    CUDA_OK(cudaStreamWaitEvent(s, c->ev_join[m - 1], 0));
  }
}

// bias gradients that no GEMM epilogue produces run on a side stream next to the fused segment that follows
void fork_colsums(Ctx* c, std::vector<Op>& ops, cudaStream_t s) {
  if (ops.empty()) return;
  CUDA_OK(cudaEventRecord(c->ev_wfork[0], s));
  CUDA_OK(cudaStreamWaitEvent(c->wstream[0], c->ev_wfork[0], 0));
  run_ops(c, ops, c->wstream[0]);
  CUDA_OK(cudaEventRecord(c->ev_wjoin[0], c->wstream[0]));
}
void join_colsums(Ctx* c, std::vector<Op>& ops, cudaStream_t s) {
  if (!ops.empty()) CUDA_OK(cudaStreamWaitEvent(s, c->ev_wjoin[0], 0));
}

// segment A1: zero grads, forward, losses, decoder backward   (gradient bucket 0 complete at its end)
void enqueue_a1(Ctx* c, cudaStream_t s, bool with_memset = true) {
  const int M = c->cfg.n_modalities;
  c->lat_mode_elt = false;
  if (one_mode(c)) {
    // ONE launch of the persistent tile kernel carries the gradient step (enqueue_a2); the gradient buffer it accumulates
    // into (TMA reduce-add, bias-gradient REDs) is cleared here -- unless the previous step's Adam left it cleared
    c->lat_mode_elt = true;
    if (with_memset) CUDA_OK(cudaMemsetAsync(c->g, 0, (size_t)(c->n_flat + 32) * sizeof(float), s));
    return;
  }
  if (elt_mode(c)) {
    // two launches of the persistent tile kernel carry the step: forward (encoders, latent stage, decoders) here,
    // backward in enqueue_a2; the gradient memset runs as a parallel branch of the forward
    c->lat_mode_elt = true;
    CUDA_OK(cudaEventRecord(c->ev_aux_fork, s));
    CUDA_OK(cudaStreamWaitEvent(c->aux_stream, c->ev_aux_fork, 0));
    CUDA_OK(cudaMemsetAsync(c->g, 0, (size_t)(c->n_flat + 32) * sizeof(float), c->aux_stream));
    CUDA_OK(cudaEventRecord(c->ev_aux_join, c->aux_stream));
    launch_seg(c, c->seg_fwd, s);
    fork_modalities(c, s);
    for (int m = 0; m < M; ++m) run_ops(c, c->ops_loss_mod[m], mod_stream(c, m, s));
    join_modalities(c, s);
    CUDA_OK(cudaStreamWaitEvent(s, c->ev_aux_join, 0));
    return;
  }
  if (c->fused) {
    // the gradient buffer is first touched by the decoder backward: its memset runs as a parallel branch of the forward
    CUDA_OK(cudaEventRecord(c->ev_aux_fork, s));
    CUDA_OK(cudaStreamWaitEvent(c->aux_stream, c->ev_aux_fork, 0));
    CUDA_OK(cudaMemsetAsync(c->g, 0, (size_t)(c->n_flat + 32) * sizeof(float), c->aux_stream));
    CUDA_OK(cudaEventRecord(c->ev_aux_join, c->aux_stream));
    // four launches of the persistent tile kernel carry every dense layer of both modalities (csrc/gemm_group.cu)
    launch_seg(c, c->seg_enc, s);
    run_ops(c, c->ops_latent_fwd, s);
    launch_seg(c, c->seg_dec, s);
    fork_modalities(c, s);
    for (int m = 0; m < M; ++m) run_ops(c, c->ops_loss_mod[m], mod_stream(c, m, s));
    join_modalities(c, s);
    CUDA_OK(cudaStreamWaitEvent(s, c->ev_aux_join, 0));
    fork_colsums(c, c->ops_colsum_dec, s);
    launch_seg(c, c->seg_bwd_dec, s);
    join_colsums(c, c->ops_colsum_dec, s);
    return;
  }
  CUDA_OK(cudaMemsetAsync(c->g, 0, (size_t)(c->n_flat + 32) * sizeof(float), s));
  fork_modalities(c, s);
  for (int m = 0; m < M; ++m) run_ops(c, c->ops_enc_mod[m], mod_stream(c, m, s));
  join_modalities(c, s);
  run_ops(c, c->ops_latent_fwd, s);
  fork_modalities(c, s);
  for (int m = 0; m < M; ++m) {
    run_ops(c, c->ops_dec_mod[m], mod_stream(c, m, s));
    run_ops(c, c->ops_loss_mod[m], mod_stream(c, m, s));
    run_bwd_ops(c, m, c->ops_bwd_dec_mod[m], mod_stream(c, m, s));
  }
  join_modalities(c, s);
}
// segment A2: latent + encoder backward, cost finalize (bucket 1 + cost slot complete at its end)
void enqueue_a2(Ctx* c, cudaStream_t s, int advance) {
  if (one_mode(c)) {
    c->lat_mode_elt = true;
    launch_seg(c, c->seg_step, s, advance);
    return;
  }
  if (elt_mode(c)) {
    c->lat_mode_elt = true;
    CUDA_OK(cudaEventRecord(c->ev_aux_fork, s));
    CUDA_OK(cudaStreamWaitEvent(c->aux_stream, c->ev_aux_fork, 0));
    launch_finalize(finalize_args(c, advance), c->aux_stream);
    c->launches += 1;
    CUDA_OK(cudaEventRecord(c->ev_aux_join, c->aux_stream));
    fork_colsums(c, c->ops_colsum_dec, s);
    launch_seg(c, c->seg_bwd, s);
    join_colsums(c, c->ops_colsum_dec, s);
    CUDA_OK(cudaStreamWaitEvent(s, c->ev_aux_join, 0));
    return;
  }
  run_ops(c, c->ops_latent_bwd, s);
  if (c->fused) {
    // the cost reduction (block partials of the loss kernels -> cost slot, step counter) runs next to the encoder backward
    CUDA_OK(cudaEventRecord(c->ev_aux_fork, s));
    CUDA_OK(cudaStreamWaitEvent(c->aux_stream, c->ev_aux_fork, 0));
    launch_finalize(finalize_args(c, advance), c->aux_stream);
    c->launches += 1;
    CUDA_OK(cudaEventRecord(c->ev_aux_join, c->aux_stream));
    fork_colsums(c, c->ops_colsum_enc, s);
    launch_seg(c, c->seg_bwd_enc, s);
    join_colsums(c, c->ops_colsum_enc, s);
    CUDA_OK(cudaStreamWaitEvent(s, c->ev_aux_join, 0));
    return;
  } else {
    fork_modalities(c, s);
    for (int m = 0; m < c->cfg.n_modalities; ++m) run_bwd_ops(c, m, c->ops_bwd_enc_mod[m], mod_stream(c, m, s));
    join_modalities(c, s);
  }
  launch_finalize(finalize_args(c, advance), s);
  c->launches += 1;
}
void enqueue_adam(Ctx* c, cudaStream_t s, bool zero_g = false) {
  AdamArgs a = adam_args(c);
  if (zero_g) a.zero_g = c->g;
  launch_adam(a, s);
  c->launches += 1;
}
// rank r owns float4 indices [lo, hi) of the flat buffers: equal shards, multiples of 8 float4 (128 bytes)
void peer_shard(const Ctx* c, int r, int64_t* lo, int64_t* hi) {
  const int64_t n4 = c->n_flat >> 2;
  const int64_t per = round_up((n4 + c->world - 1) / c->world, 8);
  *lo = std::min<int64_t>(n4, per * r);
  *hi = std::min<int64_t>(n4, per * (r + 1));
}
void enqueue_peer_adam(Ctx* c, cudaStream_t s) {
  PeerAdamArgs a;
  a.adam = adam_args(c);
  a.world = c->world; a.rank = c->rank;
  const bool shadow = a.adam.p_tf32 != nullptr;
  for (int r = 0; r < c->world; ++r) {
    float* b = c->peer.base[r];
    a.p_peer[r] = b + (c->p - c->arena);
    a.g_peer[r] = b + (c->g - c->arena);
    a.ptf_peer[r] = shadow ? b + (c->p_tf32 - c->arena) : nullptr;
    a.flag_peer[r] = reinterpret_cast<uint32_t*>(b + (reinterpret_cast<float*>(c->peer.flags) - c->arena));
  }
  if (c->peer.mc) {
    a.g_mc = c->peer.mc + (c->g - c->arena);
    a.p_mc = c->peer.mc + (c->p - c->arena);
    a.ptf_mc = shadow ? c->peer.mc + (c->p_tf32 - c->arena) : nullptr;
  }
  a.sync = c->peer.flags + 2 * kMaxPeers;
  peer_shard(c, c->rank, &a.shard_lo, &a.shard_hi);
  a.tl = c->peer.tl;
  launch_peer_adam(a, s);
  c->launches += 1;
}
// every rank's parameter stores of the previous peer step have landed here, and every rank has read our gradients
void enqueue_peer_wait(Ctx* c, cudaStream_t s) {
  launch_peer_wait(c->peer.flags, c->peer.flags + 2 * kMaxPeers, c->world, s);
  c->launches += 1;
}
void allreduce(Ctx* c, float* buf, int64_t count, cudaStream_t s);
void enqueue_forward_loss(Ctx* c, cudaStream_t s) {   // evaluate_cost: no gradients
  c->lat_mode_elt = false;
  c->recon_stale = 0;
  run_ops(c, c->ops_fwd_enc, s);
  {
    // latent forward without the gradient stash
    LatentArgs a;
    const int M = c->cfg.n_modalities;
    a.n_mod = M; a.batch = c->cfg.batch_size; a.n_z = c->cfg.n_z;
    a.inv_global_batch = 1.0f / (float)global_batch(c); a.lambda = c->cfg.assoc_lambda;
    for (int m = 0; m < M; ++m) {
      Mod& d = c->mods[m];
      a.weight[m] = d.cfg.weight; a.heads[m] = d.hd; a.z[m] = d.z; a.gstat[m] = d.gstat; a.latent_loss[m] = d.lat_loss;
    }
    a.eps = c->eps; a.partials = c->lat_partials; a.with_grad = 0; a.round_z = c->round_z ? 1 : 0;
    launch_latent_fwd(a, s);
    c->launches += 1;
  }
  run_ops(c, c->ops_fwd_dec, s);
  run_ops(c, c->ops_loss, s);   // also writes d cost/d a (harmless; gradients are not consumed)
  launch_finalize(finalize_args(c, 0), s);
  if (c->comm && c->world > 1) { allreduce(c, c->g + c->n_flat, 32, s); c->launches += 1; }
  launch_publish_cost(c->g + c->n_flat, c->last_cost, s);
  c->launches += 2;
}

// capture `fn` into an executable graph; returns the number of kernel/memset nodes
template <typename F>
int capture(Ctx* c, cudaGraphExec_t* out, F&& fn) {
  cudaStream_t s = c->own_stream;
  const int64_t launches_before = c->launches;
  CUDA_OK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  try {
    fn(s);
  } catch (...) {
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(s, &g);
    if (g) cudaGraphDestroy(g);
    c->launches = launches_before;
    throw;
  }
  cudaGraph_t g = nullptr;
  CUDA_OK(cudaStreamEndCapture(s, &g));
  const int n = (int)(c->launches - launches_before);
  c->launches = launches_before;
  cudaError_t e = cudaGraphInstantiate(out, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) fail("cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
  return n;
}

void ensure_graphs(Ctx* c) {
  if (!c->cfg.use_graph || c->graph_a1) return;
  c->graph_a1_nodes = capture(c, &c->graph_a1, [&](cudaStream_t s) { enqueue_a1(c, s); });
  c->graph_a2_nodes = capture(c, &c->graph_a2, [&](cudaStream_t s) { enqueue_a2(c, s, 1); });
  c->graph_adam_nodes = capture(c, &c->graph_adam, [&](cudaStream_t s) { enqueue_adam(c, s); });
  c->graph_train_nodes = capture(c, &c->graph_train, [&](cudaStream_t s) {
    enqueue_a1(c, s); enqueue_a2(c, s, 1); enqueue_adam(c, s);
  });
  if (one_mode(c))
    c->graph_train_nz_nodes = capture(c, &c->graph_train_nz, [&](cudaStream_t s) {
      enqueue_a1(c, s, false); enqueue_a2(c, s, 1); enqueue_adam(c, s, true);
    });
  c->graph_grad_nodes = capture(c, &c->graph_grad, [&](cudaStream_t s) {
    enqueue_a1(c, s); enqueue_a2(c, s, 0);
    launch_publish_cost(c->g + c->n_flat, c->last_cost, s); c->launches += 1;
  });
}

void refresh_shadow(Ctx* c, cudaStream_t s) {
  if (c->shadow_dirty && c->cfg.precision == VAEASSOC_TF32) {
    launch_round_copy(c->p, c->p_tf32, c->n_flat, s);
    c->launches += 1;
  }
  c->shadow_dirty = false;
}

void stage_inputs(Ctx* c, const float* const* x, const int64_t* ld, const float* eps, cudaStream_t s,
                  int only_modality = -1, bool want_eps = true, const int64_t* row_index = nullptr) {
  StageArgs a;
  if (only_modality < 0 && c->split_heads && one_mode(c)) {     // a train / gradient step follows
    int z = 0;
    for (int m = 0; m < c->cfg.n_modalities && z + 1 < 8; ++m) {
      const Mod& d = c->mods[m];
      a.zero_ptr[z] = d.hd; a.zero_n[z++] = (int64_t)c->cfg.batch_size * d.nh;
      a.zero_ptr[z] = d.dz; a.zero_n[z++] = (int64_t)c->cfg.batch_size * c->cfg.n_z;
    }
  }
  a.row_index = row_index;
  a.n_mod = c->cfg.n_modalities; a.batch = c->cfg.batch_size;
  for (int m = 0; m < a.n_mod; ++m) {
    const Mod& d = c->mods[m];
    const bool use = (only_modality < 0 || only_modality == m) && x && x[m];
    a.src[m] = use ? x[m] : nullptr;
    a.src_ld[m] = (ld && ld[m] > 0) ? ld[m] : d.ni;
    a.dst[m] = d.xs; a.dst_ld[m] = d.nip; a.n_input[m] = d.ni;
  }
  a.round_tf32 = c->round_x ? 1 : 0;
  a.eps_src = eps; a.eps_dst = want_eps ? c->eps : nullptr; a.n_z = c->cfg.n_z;
  a.eps_seed = c->cfg.eps_seed; a.global_row0 = c->cfg.global_row0; a.step_dev = c->step_dev;
  launch_stage(a, s);
  c->launches += 1;
}

// ncclCommGetAsyncError -> abort: a failed peer / link must surface as an error of the next call, not as a hang
void comm_check(Ctx* c) {
  if (!c->comm) return;
  int async_err = 0;
  const int r = g_nccl.CommGetAsyncError(c->comm, &async_err);
  if (r != 0 || async_err != 0) {
    const char* what = g_nccl.GetErrorString(r != 0 ? r : async_err);
    g_nccl.CommAbort(c->comm);
    c->comm = nullptr; c->world = 1; c->rank = 0;
    destroy_graphs(c);
    fail("NCCL asynchronous error: %s -- communicator aborted", what ? what : "unknown");
  }
}

void allreduce(Ctx* c, float* buf, int64_t count, cudaStream_t s) {
  const int r = g_nccl.AllReduce(buf, buf, (size_t)count, kNcclFloat, kNcclSum, c->comm, s);
  if (r != 0) fail("ncclAllReduce failed: %s", g_nccl.GetErrorString(r));
  if ((++c->comm_calls & 63) == 0) comm_check(c);       // cheap host-side poll, every 64 collectives
}

// the train step proper (inputs already staged)
void run_step(Ctx* c, bool with_adam) {
  cudaStream_t s = c->stream;
  refresh_shadow(c, s);
  ensure_graphs(c);
  if (one_mode(c)) c->recon_stale = (1u << c->cfg.n_modalities) - 1u;
  const bool dp = c->comm != nullptr && c->world > 1;
  if (dp && c->peer.on && with_adam) {
    // peer-memory step: forward, backward and ONE kernel that reduce-scatters the gradients over NVLink, runs Adam on
    // the owned shard and all-gathers the parameters (peer_adam.cu) -- no NCCL, one graph launch per step
    c->peer.slots_stale = true;
    if (c->cfg.use_graph) {
      if (!c->peer.graph)
        c->peer.graph_nodes = capture(c, &c->peer.graph, [&](cudaStream_t cs) {
          enqueue_peer_wait(c, cs); enqueue_a1(c, cs); enqueue_a2(c, cs, 1); enqueue_peer_adam(c, cs);
        });
      CUDA_OK(cudaGraphLaunch(c->peer.graph, s));
      c->launches += c->peer.graph_nodes;
      c->peer.pending = true;
    } else {
      enqueue_peer_wait(c, s);
      c->peer.pending = true;
      enqueue_a1(c, s);
      enqueue_a2(c, s, 1);
      enqueue_peer_adam(c, s);
    }
    return;
  }
  const bool was_zero = c->g_zero;
  c->g_zero = false;
  if (!dp) {
    if (c->cfg.use_graph && with_adam && c->graph_train_nz && !getenv("VAEASSOC_KEEP_GRADS")) {
      // the train graph without a memset: Adam clears the gradients it has consumed (get_grads() after a train step
      // therefore reads zeros; compute_gradients keeps them)
      if (!was_zero) CUDA_OK(cudaMemsetAsync(c->g, 0, (size_t)(c->n_flat + 32) * sizeof(float), s));
      CUDA_OK(cudaGraphLaunch(c->graph_train_nz, s));
      c->launches += c->graph_train_nz_nodes;
      c->g_zero = true;
    } else if (c->cfg.use_graph) {
      CUDA_OK(cudaGraphLaunch(with_adam ? c->graph_train : c->graph_grad, s));
      c->launches += with_adam ? c->graph_train_nodes : c->graph_grad_nodes;
    } else {
      enqueue_a1(c, s);
      enqueue_a2(c, s, with_adam ? 1 : 0);
      if (with_adam) enqueue_adam(c, s);
      else { launch_publish_cost(c->g + c->n_flat, c->last_cost, s); c->launches += 1; }
    }
    return;
  }
  if (c->dp_single >= 0 ? c->dp_single != 0 : (c->fused && c->world > 4)) {
    // one all-reduce of the whole flat gradient buffer (+ cost slot) on the compute stream after the backward pass: half
    // the NCCL launches and no cross-stream events.  The persistent tile kernels occupy every SM, so the bucket-0
    // all-reduce of the two-bucket schedule below cannot run next to the encoder backward; its 24 NVLS CTAs only take
    // SMs away from the tile kernel while they wait for the slowest rank.  Measured at 8 ranks, 8192 pairs per rank:
    // 0.413 ms per step against 0.653 ms with the two overlapped buckets (0.320 ms on one GPU).  With 2 and 4 ranks the
    // overlapped buckets are still slightly ahead (0.366 / 0.374 ms against 0.380 at 2 ranks), hence world > 4
    if (c->cfg.use_graph) { CUDA_OK(cudaGraphLaunch(c->graph_a1, s)); c->launches += c->graph_a1_nodes; }
    else enqueue_a1(c, s);
    if (c->cfg.use_graph && with_adam) { CUDA_OK(cudaGraphLaunch(c->graph_a2, s)); c->launches += c->graph_a2_nodes; }
    else enqueue_a2(c, s, with_adam ? 1 : 0);
    allreduce(c, c->g, c->n_flat + 32, s);
    c->launches += 1;
    if (with_adam) {
      if (c->cfg.use_graph) { CUDA_OK(cudaGraphLaunch(c->graph_adam, s)); c->launches += c->graph_adam_nodes; }
      else enqueue_adam(c, s);
    } else {
      launch_publish_cost(c->g + c->n_flat, c->last_cost, s);
      c->launches += 1;
    }
    return;
  }
  // data parallel: bucket 0 (decoders) is all-reduced on the comm stream while the encoder backward runs
  // (capturing the NCCL kernels into one whole-step graph was tried and hung at communicator teardown / on the
  // synchronous partial_fit path with NCCL 2.28.9: the collectives stay eager launches between three graphs)
  if (c->cfg.use_graph) { CUDA_OK(cudaGraphLaunch(c->graph_a1, s)); c->launches += c->graph_a1_nodes; }
  else enqueue_a1(c, s);
  CUDA_OK(cudaEventRecord(c->ev_bucket, s));
  CUDA_OK(cudaStreamWaitEvent(c->comm_stream, c->ev_bucket, 0));
  allreduce(c, c->g, c->bucket_split, c->comm_stream);
  if (c->cfg.use_graph && with_adam) { CUDA_OK(cudaGraphLaunch(c->graph_a2, s)); c->launches += c->graph_a2_nodes; }
  else enqueue_a2(c, s, with_adam ? 1 : 0);
  CUDA_OK(cudaEventRecord(c->ev_bucket, s));
  CUDA_OK(cudaStreamWaitEvent(c->comm_stream, c->ev_bucket, 0));
  allreduce(c, c->g + c->bucket_split, c->n_flat + 32 - c->bucket_split, c->comm_stream);
  CUDA_OK(cudaEventRecord(c->ev_comm, c->comm_stream));
  CUDA_OK(cudaStreamWaitEvent(s, c->ev_comm, 0));
  if (with_adam) {
    if (c->cfg.use_graph) { CUDA_OK(cudaGraphLaunch(c->graph_adam, s)); c->launches += c->graph_adam_nodes; }
    else enqueue_adam(c, s);
  } else {
    launch_publish_cost(c->g + c->n_flat, c->last_cost, s);
    c->launches += 1;
  }
}

void check_handle(vaeassoc_handle h) {
  if (!h) fail("null handle");
  CUDA_OK(cudaSetDevice(h->device));
}

void check_precision(int precision) {
  if (precision == VAEASSOC_BF16)
    fail("precision bf16 is not served: 8-bit mantissa operands miss the 2e-3 tolerance of the tensor-core path "
         "(tf32 measures 6e-4); use tf32 (tcgen05 kind::tf32) or fp32");
  if (precision != VAEASSOC_FP32 && precision != VAEASSOC_TF32) fail("precision must be fp32 or tf32");
}

void validate_config(const vaeassoc_config* cfg) {
  if (!cfg) fail("null config");
  if (cfg->abi_version != VAEASSOC_ABI_VERSION) fail("ABI version mismatch: caller %d, library %d", cfg->abi_version, VAEASSOC_ABI_VERSION);
  if (cfg->n_modalities < 1 || cfg->n_modalities > VAEASSOC_MAX_MODALITIES) fail("n_modalities must be in [1,%d]", VAEASSOC_MAX_MODALITIES);
  if (cfg->batch_size < 1) fail("batch_size must be >= 1");
  if (cfg->n_z < 1 || cfg->n_z > 1024) fail("n_z must be in [1,1024]");
  if (cfg->transfer_fct != VAEASSOC_RELU && cfg->transfer_fct != VAEASSOC_SOFTPLUS) fail("transfer_fct must be relu or softplus");
  check_precision(cfg->precision);
  if (cfg->global_batch != 0 && cfg->global_batch < cfg->batch_size) fail("global_batch < batch_size");
}

}  // namespace

// =======================================================================================================
// C-ABI
// =======================================================================================================
// (every entry point first makes the peers' parameter stores of the last data-parallel step visible -- peer_quiesce --
// except the training entry points, whose step graph starts with that wait so that it overlaps the input staging)
#define API_BEGIN_TRAIN(h)                    \
  try {                                       \
    check_handle(h);                          \
    std::lock_guard<std::mutex> lock__(h->mu);
#define API_BEGIN(h)                          \
  API_BEGIN_TRAIN(h)                          \
    peer_quiesce(h);
#define API_END(h)                            \
    return 0;                                 \
  } catch (const std::exception& e) {         \
    if (h) h->err = e.what(); else g_create_error = e.what(); \
    return 1;                                 \
  }

extern "C" {

int vaeassoc_abi_version(void) { return VAEASSOC_ABI_VERSION; }

const char* vaeassoc_last_error(vaeassoc_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int vaeassoc_create(const vaeassoc_config* cfg, vaeassoc_handle* out) {
  Ctx* c = nullptr;
  try {
    if (!out) fail("null output handle");
    *out = nullptr;
    validate_config(cfg);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      fail("no CUDA device available (%s); libvaeassoc has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) fail("device %d out of range (have %d)", cfg->device, ndev);
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) fail("device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor);
    CUDA_OK(cudaSetDevice(cfg->device));
    c = new Ctx();
    c->cfg = *cfg;
    c->device = cfg->device;
    if (c->cfg.beta1 == 0.f) c->cfg.beta1 = 0.9f;
    if (c->cfg.beta2 == 0.f) c->cfg.beta2 = 0.999f;
    if (c->cfg.adam_epsilon == 0.f) c->cfg.adam_epsilon = 1e-8f;
    CUDA_OK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    c->force_dynamic = getenv("VAEASSOC_DYNAMIC_FIRST") != nullptr;
    if (const char* e = getenv("VAEASSOC_DP_SINGLE")) c->dp_single = atoi(e) != 0 ? 1 : 0;
    for (int i = 0; i < 2; ++i) {
      CUDA_OK(cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < Ctx::kUploadRing; ++i) CUDA_OK(cudaEventCreateWithFlags(&c->ev_upload[i], cudaEventDisableTiming));
    for (int i = 0; i < VAEASSOC_MAX_MODALITIES - 1; ++i) {
      CUDA_OK(cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < VAEASSOC_MAX_MODALITIES; ++i) {
      CUDA_OK(cudaStreamCreateWithFlags(&c->wstream[i], cudaStreamNonBlocking));
      CUDA_OK(cudaEventCreateWithFlags(&c->ev_wfork[i], cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&c->ev_wjoin[i], cudaEventDisableTiming));
    }
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUDA_OK(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_aux_fork, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_aux_join, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_bucket, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming));
    build_layout(c);
    alloc_buffers(c);
    build_ops(c);
    CUDA_OK(cudaDeviceSynchronize());
    *out = c;
    return 0;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    if (c) vaeassoc_destroy(c);
    return 1;
  }
}

int vaeassoc_destroy(vaeassoc_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  peer_detach(h);
  if (h->comm) { g_nccl.CommDestroy(h->comm); h->comm = nullptr; }
  destroy_graphs(h);
  group_destroy(h->gplan);
  for (void* p : h->allocs) cudaFree(p);
  h->free_op_ws();
  if (h->host_cost_ring) cudaFreeHost(h->host_cost_ring);
  if (h->inf.pin_in) cudaFreeHost(h->inf.pin_in);
  if (h->inf.pin_out) cudaFreeHost(h->inf.pin_out);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
    if (h->ev_consumed[i]) cudaEventDestroy(h->ev_consumed[i]);
  }
  for (int i = 0; i < Ctx::kUploadRing; ++i) if (h->ev_upload[i]) cudaEventDestroy(h->ev_upload[i]);
  for (int i = 0; i < VAEASSOC_MAX_MODALITIES - 1; ++i) {
    if (h->side[i]) cudaStreamDestroy(h->side[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
  }
  for (int i = 0; i < VAEASSOC_MAX_MODALITIES; ++i) {
    if (h->wstream[i]) cudaStreamDestroy(h->wstream[i]);
    if (h->ev_wfork[i]) cudaEventDestroy(h->ev_wfork[i]);
    if (h->ev_wjoin[i]) cudaEventDestroy(h->ev_wjoin[i]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  if (h->ev_aux_fork) cudaEventDestroy(h->ev_aux_fork);
  if (h->ev_aux_join) cudaEventDestroy(h->ev_aux_join);
  if (h->ev_bucket) cudaEventDestroy(h->ev_bucket);
  if (h->ev_comm) cudaEventDestroy(h->ev_comm);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  delete h;
  return 0;
}

int vaeassoc_set_stream(vaeassoc_handle h, void* cuda_stream) {
  API_BEGIN(h)
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  API_END(h)
}

int vaeassoc_stream_sync(vaeassoc_handle h) {
  API_BEGIN(h)
  CUDA_OK(cudaStreamSynchronize(h->stream));
  comm_check(h);
  API_END(h)
}

int vaeassoc_set_precision(vaeassoc_handle h, int precision) {
  API_BEGIN(h)
  check_precision(precision);
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->cfg.precision = precision;
  h->shadow_dirty = true;
  build_ops(h);
  API_END(h)
}

int vaeassoc_set_learning_rate(vaeassoc_handle h, float lr) {
  API_BEGIN(h)
  CUDA_OK(cudaStreamSynchronize(h->stream));
  h->cfg.learning_rate = lr;
  destroy_graphs(h);
  API_END(h)
}

int vaeassoc_num_tensors(vaeassoc_handle h) { return h ? (int)h->tensors.size() : -1; }

int vaeassoc_layout_query(vaeassoc_handle h, int i, vaeassoc_tensor_info* out) {
  API_BEGIN(h)
  if (i < 0 || i >= (int)h->tensors.size() || !out) fail("tensor index %d out of range", i);
  *out = h->tensors[i];
  API_END(h)
}

int64_t vaeassoc_flat_size(vaeassoc_handle h) { return h ? h->n_flat : -1; }

void* vaeassoc_flat_ptr(vaeassoc_handle h, int which) {
  if (!h) return nullptr;
  switch (which) {
    case VAEASSOC_PARAMS: return h->p;
    case VAEASSOC_GRADS: h->g_zero = false; return h->g;      // (the caller may write through the pointer)
    case VAEASSOC_ADAM_M: return h->m;
    case VAEASSOC_ADAM_V: return h->v;
    default: return nullptr;
  }
}

int vaeassoc_tensor_set(vaeassoc_handle h, int which, int i, const float* src_host) {
  API_BEGIN(h)
  float* base = reinterpret_cast<float*>(vaeassoc_flat_ptr(h, which));
  if (!base) fail("bad flat buffer id %d", which);
  if (i < 0 || i >= (int)h->tensors.size() || !src_host) fail("tensor index %d out of range", i);
  const vaeassoc_tensor_info& t = h->tensors[i];
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaMemcpy2D(base + t.offset, (size_t)t.ld * 4, src_host, (size_t)t.cols * 4, (size_t)t.cols * 4, (size_t)t.rows,
                       cudaMemcpyHostToDevice));
  if (which == VAEASSOC_PARAMS) h->shadow_dirty = true;
  API_END(h)
}

int vaeassoc_tensor_get(vaeassoc_handle h, int which, int i, float* dst_host) {
  API_BEGIN(h)
  float* base = reinterpret_cast<float*>(vaeassoc_flat_ptr(h, which));
  if (!base) fail("bad flat buffer id %d", which);
  if (i < 0 || i >= (int)h->tensors.size() || !dst_host) fail("tensor index %d out of range", i);
  const vaeassoc_tensor_info& t = h->tensors[i];
  CUDA_OK(cudaStreamSynchronize(h->stream));
  if (which == VAEASSOC_ADAM_M || which == VAEASSOC_ADAM_V) refresh_peer_slots(h);
  CUDA_OK(cudaMemcpy2D(dst_host, (size_t)t.cols * 4, base + t.offset, (size_t)t.ld * 4, (size_t)t.cols * 4, (size_t)t.rows,
                       cudaMemcpyDeviceToHost));
  API_END(h)
}

int vaeassoc_step_get(vaeassoc_handle h, int64_t* step) {
  API_BEGIN(h)
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaMemcpy(step, h->step_dev, sizeof(int64_t), cudaMemcpyDeviceToHost));
  API_END(h)
}

int vaeassoc_step_set(vaeassoc_handle h, int64_t step) {
  API_BEGIN(h)
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaMemcpy(h->step_dev, &step, sizeof(int64_t), cudaMemcpyHostToDevice));
  API_END(h)
}

int vaeassoc_train_step(vaeassoc_handle h, const float* const* x_dev, const int64_t* ld, const float* eps_dev) {
  API_BEGIN_TRAIN(h)
  if (!x_dev) fail("x_dev is null");
  for (int m = 0; m < h->cfg.n_modalities; ++m) if (!x_dev[m]) fail("x_dev[%d] is null", m);
  stage_inputs(h, x_dev, ld, eps_dev, h->stream);
  run_step(h, true);
  API_END(h)
}

int vaeassoc_grad_step(vaeassoc_handle h, const float* const* x_dev, const int64_t* ld, const float* eps_dev) {
  API_BEGIN(h)
  if (!x_dev) fail("x_dev is null");
  for (int m = 0; m < h->cfg.n_modalities; ++m) if (!x_dev[m]) fail("x_dev[%d] is null", m);
  stage_inputs(h, x_dev, ld, eps_dev, h->stream);
  run_step(h, false);
  API_END(h)
}

int vaeassoc_adam_step(vaeassoc_handle h) {
  API_BEGIN(h)
  // grad_step leaves t unchanged; ApplyAdam uses t+1
  int64_t one = 0;
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaMemcpy(&one, h->step_dev, sizeof one, cudaMemcpyDeviceToHost));
  one += 1;
  CUDA_OK(cudaMemcpy(h->step_dev, &one, sizeof one, cudaMemcpyHostToDevice));
  enqueue_adam(h, h->stream);
  API_END(h)
}

int vaeassoc_cost_read(vaeassoc_handle h, float* cost_host) {
  API_BEGIN(h)
  CUDA_OK(cudaMemcpyAsync(cost_host, h->last_cost, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  API_END(h)
}

int vaeassoc_cost_history(vaeassoc_handle h, int64_t first_step, int64_t n, float* dst_host) {
  API_BEGIN(h)
  if (n < 0 || n > h->hist_cap) fail("history window %lld exceeds the ring capacity %d", (long long)n, h->hist_cap);
  CUDA_OK(cudaStreamSynchronize(h->stream));
  for (int64_t i = 0; i < n;) {
    const int64_t pos = (first_step + i) % h->hist_cap;
    const int64_t run = std::min<int64_t>(n - i, h->hist_cap - pos);
    CUDA_OK(cudaMemcpy(dst_host + i, h->cost_hist + pos, (size_t)run * 4, cudaMemcpyDeviceToHost));
    i += run;
  }
  API_END(h)
}

static void upload_host(Ctx* h, const float* const* x_host, const float* eps_host, int slot, cudaStream_t s) {
  for (int m = 0; m < h->cfg.n_modalities; ++m) {
    if (!x_host[m]) fail("x_host[%d] is null", m);
    const Mod& d = h->mods[m];
    CUDA_OK(cudaMemcpyAsync(d.xin[slot], x_host[m], (size_t)h->cfg.batch_size * d.ni * 4, cudaMemcpyHostToDevice, s));
  }
  if (eps_host)
    CUDA_OK(cudaMemcpyAsync(h->eps_in[slot], eps_host, (size_t)h->cfg.batch_size * h->cfg.n_z * 4, cudaMemcpyHostToDevice, s));
}

int vaeassoc_partial_fit_host(vaeassoc_handle h, const float* const* x_host, const float* eps_host, float* cost_host) {
  API_BEGIN_TRAIN(h)
  if (!x_host) fail("x_host is null");
  upload_host(h, x_host, eps_host, 0, h->stream);
  const float* xd[VAEASSOC_MAX_MODALITIES];
  for (int m = 0; m < h->cfg.n_modalities; ++m) xd[m] = h->mods[m].xin[0];
  stage_inputs(h, xd, nullptr, eps_host ? h->eps_in[0] : nullptr, h->stream);
  run_step(h, true);
  if (cost_host) CUDA_OK(cudaMemcpyAsync(cost_host, h->last_cost, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  API_END(h)
}

int vaeassoc_submit_host(vaeassoc_handle h, const float* const* x_host, const float* eps_host) {
  API_BEGIN_TRAIN(h)
  if (!x_host) fail("x_host is null");
  const int slot = (int)(h->submit_count & 1);
  // the upload of batch k+1 overlaps the compute of batch k; slot reuse waits for the stage kernel of batch k-1
  if (h->submit_count >= 2) CUDA_OK(cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[slot], 0));
  upload_host(h, x_host, eps_host, slot, h->copy_stream);
  CUDA_OK(cudaEventRecord(h->ev_copied[slot], h->copy_stream));
  CUDA_OK(cudaEventRecord(h->ev_upload[h->submit_count % Ctx::kUploadRing], h->copy_stream));
  CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_copied[slot], 0));
  const float* xd[VAEASSOC_MAX_MODALITIES];
  for (int m = 0; m < h->cfg.n_modalities; ++m) xd[m] = h->mods[m].xin[slot];
  stage_inputs(h, xd, nullptr, eps_host ? h->eps_in[slot] : nullptr, h->stream);
  CUDA_OK(cudaEventRecord(h->ev_consumed[slot], h->stream));
  run_step(h, true);
  CUDA_OK(cudaMemcpyAsync(h->host_cost_ring + (h->submit_count % Ctx::kHostRing), h->last_cost, sizeof(float),
                          cudaMemcpyDeviceToHost, h->stream));
  h->submit_count += 1;
  API_END(h)
}

int vaeassoc_submit_indexed(vaeassoc_handle h, const float* data_dev, int64_t ld, int64_t n_rows,
                            const int64_t* index_host, const float* eps_host) {
  API_BEGIN_TRAIN(h)
  if (!data_dev || !index_host) fail("data_dev / index_host is null");
  int64_t width = 0;
  for (int m = 0; m < h->cfg.n_modalities; ++m) width += h->mods[m].ni;
  if (ld < width) fail("row pitch %lld < %lld columns (sum of n_input)", (long long)ld, (long long)width);
  const int64_t B = h->cfg.batch_size;
  for (int64_t r = 0; r < B; ++r)
    if (index_host[r] < 0 || index_host[r] >= n_rows) fail("row index %lld out of range [0, %lld)", (long long)index_host[r], (long long)n_rows);
  const int slot = (int)(h->indexed_count & 1);
  cudaStream_t s = h->stream;
  // 8 B per pair of indices (+ eps when injected) is all that crosses PCIe; the rows are gathered on the device from the
  // data set uploaded once.  Stream order makes the slot reuse safe (the stage kernel two submits back has run).
  CUDA_OK(cudaMemcpyAsync(h->idx_in[slot], index_host, (size_t)B * sizeof(int64_t), cudaMemcpyHostToDevice, s));
  if (eps_host) CUDA_OK(cudaMemcpyAsync(h->eps_in[slot], eps_host, (size_t)B * h->cfg.n_z * 4, cudaMemcpyHostToDevice, s));
  const float* xd[VAEASSOC_MAX_MODALITIES] = {nullptr, nullptr, nullptr, nullptr};
  int64_t lds[VAEASSOC_MAX_MODALITIES] = {0, 0, 0, 0};
  int64_t col = 0;
  for (int m = 0; m < h->cfg.n_modalities; ++m) { xd[m] = data_dev + col; lds[m] = ld; col += h->mods[m].ni; }
  stage_inputs(h, xd, lds, eps_host ? h->eps_in[slot] : nullptr, s, -1, true, h->idx_in[slot]);
  run_step(h, true);
  CUDA_OK(cudaMemcpyAsync(h->host_cost_ring + (h->submit_count % Ctx::kHostRing), h->last_cost, sizeof(float),
                          cudaMemcpyDeviceToHost, s));
  // (no host buffer of the caller is read after this call returns except index_host / eps_host when they are pinned:
  // same lifetime rule as vaeassoc_submit_host, vaeassoc_upload_wait applies)
  CUDA_OK(cudaEventRecord(h->ev_upload[h->submit_count % Ctx::kUploadRing], s));
  h->indexed_count += 1;
  h->submit_count += 1;
  API_END(h)
}

int64_t vaeassoc_submit_count(vaeassoc_handle h) { return h ? h->submit_count : -1; }

int vaeassoc_upload_wait(vaeassoc_handle h, int64_t submit_index) {
  API_BEGIN(h)
  if (submit_index < 0 || submit_index >= h->submit_count) fail("submit %lld has not been made (%lld so far)", (long long)submit_index, (long long)h->submit_count);
  // the copy stream is in order: an event of a LATER submit implies this one; the ring holds the last kUploadRing
  const int64_t oldest = std::max<int64_t>(0, h->submit_count - Ctx::kUploadRing);
  const int64_t k = std::max(submit_index, oldest);
  CUDA_OK(cudaEventSynchronize(h->ev_upload[k % Ctx::kUploadRing]));
  API_END(h)
}

int vaeassoc_submit_costs(vaeassoc_handle h, int64_t first_submit, int64_t n, float* dst_host) {
  API_BEGIN(h)
  if (n < 0 || n > Ctx::kHostRing || first_submit < 0 || first_submit + n > h->submit_count)
    fail("submit window [%lld,+%lld) is outside the %d most recent of %lld submits", (long long)first_submit,
         (long long)n, Ctx::kHostRing, (long long)h->submit_count);
  CUDA_OK(cudaStreamSynchronize(h->stream));
  for (int64_t i = 0; i < n; ++i) dst_host[i] = h->host_cost_ring[(first_submit + i) % Ctx::kHostRing];
  API_END(h)
}

int vaeassoc_eval_cost(vaeassoc_handle h, const float* const* x_dev, const int64_t* ld, const float* eps_dev,
                       float* cost_host) {
  API_BEGIN(h)
  if (!x_dev) fail("x_dev is null");
  for (int m = 0; m < h->cfg.n_modalities; ++m) if (!x_dev[m]) fail("x_dev[%d] is null", m);
  // under data parallelism every rank calls this with ITS shard: the shard costs (global scaling baked in) are summed
  stage_inputs(h, x_dev, ld, eps_dev, h->stream);
  refresh_shadow(h, h->stream);
  enqueue_forward_loss(h, h->stream);
  if (cost_host) CUDA_OK(cudaMemcpyAsync(cost_host, h->last_cost, sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CUDA_OK(cudaStreamSynchronize(h->stream));
  API_END(h)
}

int vaeassoc_encode(vaeassoc_handle h, int modality, const float* x_dev, int64_t ld, float* mu_dev, float* logvar_dev) {
  API_BEGIN(h)
  if (modality < 0 || modality >= h->cfg.n_modalities) fail("modality %d out of range", modality);
  if (!x_dev) fail("x_dev is null");
  const float* xs[VAEASSOC_MAX_MODALITIES] = {nullptr, nullptr, nullptr, nullptr};
  int64_t lds[VAEASSOC_MAX_MODALITIES] = {0, 0, 0, 0};
  xs[modality] = x_dev; lds[modality] = ld;
  stage_inputs(h, xs, lds, nullptr, h->stream, modality, /*want_eps=*/false);
  refresh_shadow(h, h->stream);
  run_ops(h, h->ops_enc_mod[modality], h->stream);
  const Mod& d = h->mods[modality];
  const size_t w = (size_t)d.nz * 4;
  if (mu_dev) CUDA_OK(cudaMemcpy2DAsync(mu_dev, w, d.hd, (size_t)d.nh * 4, w, (size_t)h->cfg.batch_size, cudaMemcpyDeviceToDevice, h->stream));
  if (logvar_dev) CUDA_OK(cudaMemcpy2DAsync(logvar_dev, w, d.hd + d.nz, (size_t)d.nh * 4, w, (size_t)h->cfg.batch_size, cudaMemcpyDeviceToDevice, h->stream));
  API_END(h)
}

int vaeassoc_decode(vaeassoc_handle h, int modality, const float* z_dev, float* xhat_dev) {
  API_BEGIN(h)
  if (modality < 0 || modality >= h->cfg.n_modalities) fail("modality %d out of range", modality);
  if (!z_dev || !xhat_dev) fail("null z_dev / xhat_dev");
  const Mod& d = h->mods[modality];
  const int64_t B = h->cfg.batch_size;
  CUDA_OK(cudaMemcpyAsync(d.z, z_dev, (size_t)B * d.nz * 4, cudaMemcpyDeviceToDevice, h->stream));
  refresh_shadow(h, h->stream);
  run_ops(h, h->ops_dec_mod[modality], h->stream);
  CUDA_OK(cudaMemcpy2DAsync(xhat_dev, (size_t)d.ni * 4, d.xh, (size_t)d.nip * 4, (size_t)d.ni * 4, (size_t)B, cudaMemcpyDeviceToDevice, h->stream));
  API_END(h)
}

int vaeassoc_reconstruct(vaeassoc_handle h, int modality, const float* x_dev, int64_t ld, const float* eps_dev,
                         float* xhat_dev) {
  API_BEGIN(h)
  if (modality < 0 || modality >= h->cfg.n_modalities) fail("modality %d out of range", modality);
  if (!x_dev || !xhat_dev) fail("null x_dev / xhat_dev");
  const float* xs[VAEASSOC_MAX_MODALITIES] = {nullptr, nullptr, nullptr, nullptr};
  int64_t lds[VAEASSOC_MAX_MODALITIES] = {0, 0, 0, 0};
  xs[modality] = x_dev; lds[modality] = ld;
  stage_inputs(h, xs, lds, eps_dev, h->stream, modality, /*want_eps=*/true);
  refresh_shadow(h, h->stream);
  run_ops(h, h->ops_enc_mod[modality], h->stream);
  const Mod& d = h->mods[modality];
  LatentArgs a;
  a.n_mod = 1; a.batch = h->cfg.batch_size; a.n_z = h->cfg.n_z; a.inv_global_batch = 0.f; a.lambda = 0.f;
  a.weight[0] = 0.f; a.heads[0] = d.hd; a.z[0] = d.z; a.gstat[0] = d.gstat; a.latent_loss[0] = nullptr;
  a.eps = h->eps; a.partials = h->lat_partials; a.with_grad = 0; a.round_z = h->round_z ? 1 : 0;
  launch_latent_fwd(a, h->stream);
  h->launches += 1;
  run_ops(h, h->ops_dec_mod[modality], h->stream);
  CUDA_OK(cudaMemcpy2DAsync(xhat_dev, (size_t)d.ni * 4, d.xh, (size_t)d.nip * 4, (size_t)d.ni * 4, (size_t)h->cfg.batch_size,
                            cudaMemcpyDeviceToDevice, h->stream));
  API_END(h)
}

int vaeassoc_probe_get(vaeassoc_handle h, int kind, int modality, float* dst_host, int64_t capacity, int64_t* n_written) {
  API_BEGIN(h)
  if (!dst_host) fail("dst_host is null");
  const bool needs_mod = !(kind == VAEASSOC_PROBE_ASSOC_COST || kind == VAEASSOC_PROBE_EPS);
  if (needs_mod && (modality < 0 || modality >= h->cfg.n_modalities)) fail("modality %d out of range", modality);
  const int64_t B = h->cfg.batch_size;
  const int nz = h->cfg.n_z;
  const Mod* d = needs_mod ? &h->mods[modality] : nullptr;
  const float* src = nullptr;
  int64_t rows = 1, cols = 1, ld = 1;
  switch (kind) {
    case VAEASSOC_PROBE_Z_MEAN: src = d->hd; rows = B; cols = nz; ld = d->nh; break;
    case VAEASSOC_PROBE_Z_LOG_SIGMA_SQ: src = d->hd + nz; rows = B; cols = nz; ld = d->nh; break;
    case VAEASSOC_PROBE_Z: src = d->z; rows = B; cols = nz; ld = nz; break;
    case VAEASSOC_PROBE_X_RECONSTR_MEAN: src = d->xh; rows = B; cols = d->ni; ld = d->nip; break;
    case VAEASSOC_PROBE_RECONSTR_LOSS:
      if (d->cfg.binary) { src = d->rec_loss; rows = 1; cols = B; ld = B; }
      else { src = h->scalars + 4 + modality; }
      break;
    case VAEASSOC_PROBE_LATENT_LOSS: src = d->lat_loss; rows = 1; cols = B; ld = B; break;
    case VAEASSOC_PROBE_VAE_COST: src = h->scalars + modality; break;
    case VAEASSOC_PROBE_ASSOC_COST: src = h->scalars + 8; break;
    case VAEASSOC_PROBE_D_Z_MEAN: src = d->dhd; rows = B; cols = nz; ld = d->nh; break;
    case VAEASSOC_PROBE_D_Z_LOG_SIGMA_SQ: src = d->dhd + nz; rows = B; cols = nz; ld = d->nh; break;
    case VAEASSOC_PROBE_EPS: src = h->eps; rows = B; cols = nz; ld = nz; break;
    default: fail("unknown probe kind %d", kind);
  }
  if (rows * cols > capacity) fail("probe needs %lld floats, capacity %lld", (long long)(rows * cols), (long long)capacity);
  if ((kind == VAEASSOC_PROBE_X_RECONSTR_MEAN || (kind == VAEASSOC_PROBE_RECONSTR_LOSS && d->cfg.binary)) &&
      (h->recon_stale >> modality) & 1u) {
    // the one-launch step turns the decoder's accumulator straight into loss and gradient: x_hat and the per-row losses
    // are produced on demand from the retained decoder activations (same batch, same eps, current weights)
    refresh_shadow(h, h->stream);
    run_ops(h, h->ops_dec_mod[modality], h->stream);     // (dec1, dec2 recomputed as well: identical values)
    run_ops(h, h->ops_loss_mod[modality], h->stream);
    h->recon_stale &= ~(1u << modality);
  }
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaMemcpy2D(dst_host, (size_t)cols * 4, src, (size_t)ld * 4, (size_t)cols * 4, (size_t)rows, cudaMemcpyDeviceToHost));
  if (n_written) *n_written = rows * cols;
  API_END(h)
}

int vaeassoc_synth_batch(vaeassoc_handle h, uint32_t data_seed, uint32_t proj_seed, int64_t row0, int64_t n_rows,
                         float* const* x_dev) {
  API_BEGIN(h)
  if (!x_dev) fail("x_dev is null");
  for (int m = 0; m < h->cfg.n_modalities; ++m) {
    Mod& d = h->mods[m];
    if (!x_dev[m]) fail("x_dev[%d] is null", m);
    if (d.proj_seed_built != proj_seed) {
      launch_synth_projection(d.P, d.inv_std, m, d.ni, proj_seed, h->stream);
      h->launches += 2;
      d.proj_seed_built = proj_seed;
    }
    launch_synth_modality(x_dev[m], d.ni, d.P, d.inv_std, m, d.ni, d.cfg.binary, data_seed, row0, n_rows, h->stream);
    h->launches += 1;
  }
  API_END(h)
}

int vaeassoc_philox_normal(vaeassoc_handle h, uint32_t seed, uint32_t tag, int64_t row0, int64_t n_rows, int32_t n_cols,
                           uint32_t step, float* dst_dev) {
  API_BEGIN(h)
  if (!dst_dev || n_rows < 0 || n_cols < 1) fail("bad arguments");
  launch_philox_normal(dst_dev, n_rows, n_cols, seed, tag, row0, step, h->stream);
  h->launches += 1;
  API_END(h)
}

int vaeassoc_comm_unique_id(const char* nccl_lib_path, void* id128) {
  try {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    g_nccl.load(nccl_lib_path);
    NcclId id;
    const int r = g_nccl.GetUniqueId(&id);
    if (r != 0) fail("ncclGetUniqueId failed: %s", g_nccl.GetErrorString(r));
    memcpy(id128, &id, sizeof id);
    return 0;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    return 1;
  }
}

int vaeassoc_comm_init(vaeassoc_handle h, const char* nccl_lib_path, const void* id128, int rank, int world) {
  API_BEGIN(h)
  if (world < 1 || rank < 0 || rank >= world) fail("bad rank %d / world %d", rank, world);
  if (h->comm) fail("communicator already initialised");
  {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    g_nccl.load(nccl_lib_path);
  }
  NcclId id;
  memcpy(&id, id128, sizeof id);
  void* comm = nullptr;
  const int r = g_nccl.CommInitRank(&comm, world, id, rank);
  if (r != 0) fail("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
  h->comm = comm; h->rank = rank; h->world = world;
  destroy_graphs(h);     // the captured launches bake in the task-queue mode (see group_launch)
  API_END(h)
}

int vaeassoc_comm_destroy(vaeassoc_handle h) {
  API_BEGIN(h)
  if (h->comm) {
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaStreamSynchronize(h->comm_stream));
    refresh_peer_slots(h);
    peer_detach(h);
    g_nccl.CommDestroy(h->comm);
    h->comm = nullptr; h->world = 1; h->rank = 0;
    destroy_graphs(h);
  }
  API_END(h)
}

int vaeassoc_comm_sync_state(vaeassoc_handle h) {
  API_BEGIN(h)
  if (h->comm && h->world > 1) {
    cudaStream_t s = h->stream;
    auto bc = [&](void* buf, size_t count, int type) {
      const int r = g_nccl.Broadcast(buf, buf, count, type, 0, h->comm, s);
      if (r != 0) fail("ncclBroadcast failed: %s", g_nccl.GetErrorString(r));
    };
    bc(h->p, (size_t)h->n_flat, kNcclFloat);
    bc(h->m, (size_t)h->n_flat, kNcclFloat);
    bc(h->v, (size_t)h->n_flat, kNcclFloat);
    bc(h->step_dev, 1, kNcclInt64);
    h->launches += 4;
    h->shadow_dirty = true;
    h->peer.slots_stale = false;
    CUDA_OK(cudaStreamSynchronize(s));
    comm_check(h);
  }
  API_END(h)
}

int vaeassoc_comm_check(vaeassoc_handle h) {
  API_BEGIN(h)
  comm_check(h);
  API_END(h)
}

int vaeassoc_probe_mask(vaeassoc_handle h, int layer, int modality, uint32_t* dst_host, int64_t capacity_words,
                        int64_t* n_written, int64_t* words_per_row) {
  API_BEGIN(h)
  if (modality < 0 || modality >= h->cfg.n_modalities) fail("modality %d out of range", modality);
  if (layer < 0 || layer > 3 || !dst_host) fail("layer must be 0..3 (h1, h2, g1, g2) and dst_host non-null");
  const Mod& d = h->mods[modality];
  if (!h->masks_in_use[modality])
    fail("modality %d keeps no relu masks (they exist for dense relu modalities on the tensor-core path, batch >= 32)", modality);
  const uint32_t* src = layer == 0 ? d.mh1 : layer == 1 ? d.mh2 : layer == 2 ? d.mg1 : d.mg2;
  const int64_t wpr = ((layer == 0 || layer == 2 ? d.r1 : d.r2) + 31) / 32;
  const int64_t n = (int64_t)h->cfg.batch_size * wpr;
  if (n > capacity_words) fail("mask needs %lld words, capacity %lld", (long long)n, (long long)capacity_words);
  CUDA_OK(cudaStreamSynchronize(h->stream));
  CUDA_OK(cudaMemcpy(dst_host, src, (size_t)n * 4, cudaMemcpyDeviceToHost));
  if (n_written) *n_written = n;
  if (words_per_row) *words_per_row = wpr;
  API_END(h)
}

}  // extern "C"

// ---- inference surface with host buffers: transform / generate / reconstruct as ONE graph launch + ONE D2H -----------
namespace {
enum { INF_TRANSFORM = 0, INF_GENERATE = 1, INF_RECONSTRUCT = 2 };

// the launches of one inference call (captured once per kind / modality set); inputs sit in pin_in, results go to pin_out
void enqueue_infer(Ctx* c, int kind, unsigned mask, bool eps_given, cudaStream_t s) {
  const int M = c->cfg.n_modalities;
  const int64_t B = c->cfg.batch_size;
  const int nz = c->cfg.n_z;
  const bool all = mask == (1u << M) - 1u;
  // 1. host -> device (pinned source: asynchronous DMA inside the graph)
  int64_t off = 0;
  const float* xs[VAEASSOC_MAX_MODALITIES] = {nullptr, nullptr, nullptr, nullptr};
  int64_t lds[VAEASSOC_MAX_MODALITIES] = {0, 0, 0, 0};
  if (kind != INF_GENERATE) {
    for (int m = 0; m < M; ++m) {
      const int64_t n = B * c->mods[m].ni;
      if (mask & (1u << m)) {
        CUDA_OK(cudaMemcpyAsync(c->inf.xin[m], c->inf.pin_in + off, (size_t)n * 4, cudaMemcpyHostToDevice, s));
        xs[m] = c->inf.xin[m]; lds[m] = c->mods[m].ni;
      }
      off += n;
    }
  } else {
    for (int m = 0; m < M; ++m) off += B * c->mods[m].ni;
  }
  const bool z_in = kind == INF_GENERATE || (kind == INF_RECONSTRUCT && eps_given);
  if (z_in) CUDA_OK(cudaMemcpyAsync(c->inf.zin, c->inf.pin_in + off, (size_t)(B * nz) * 4, cudaMemcpyHostToDevice, s));
  // 2. forward
  if (kind == INF_TRANSFORM) {
    stage_inputs(c, xs, lds, nullptr, s, -1, /*want_eps=*/false);
    if (all && c->fused) launch_seg(c, c->seg_enc, s);
    else for (int m = 0; m < M; ++m) if (mask & (1u << m)) run_ops(c, c->ops_enc_mod[m], s);
  } else if (kind == INF_GENERATE) {
    for (int m = 0; m < M; ++m)
      if (mask & (1u << m)) CUDA_OK(cudaMemcpyAsync(c->mods[m].z, c->inf.zin, (size_t)(B * nz) * 4, cudaMemcpyDeviceToDevice, s));
    if (all && c->fused) launch_seg(c, c->seg_dec, s);
    else for (int m = 0; m < M; ++m) if (mask & (1u << m)) run_ops(c, c->ops_dec_mod[m], s);
  } else {
    stage_inputs(c, xs, lds, eps_given ? c->inf.zin : nullptr, s, -1, /*want_eps=*/true);
    if (all && elt_mode(c)) {
      launch_seg(c, c->seg_fwd, s);          // encoders, latent stage and decoders of every modality: one launch
    } else {
      for (int m = 0; m < M; ++m) {
        if (!(mask & (1u << m))) continue;
        const Mod& d = c->mods[m];
        run_ops(c, c->ops_enc_mod[m], s);
        LatentArgs a;
        a.n_mod = 1; a.batch = (int)B; a.n_z = nz; a.inv_global_batch = 0.f; a.lambda = 0.f;
        a.weight[0] = 0.f; a.heads[0] = d.hd; a.z[0] = d.z; a.gstat[0] = d.gstat; a.latent_loss[0] = nullptr;
        a.eps = c->eps; a.partials = c->lat_partials; a.with_grad = 0; a.round_z = c->round_z ? 1 : 0;
        launch_latent_fwd(a, s);
        c->launches += 1;
        run_ops(c, c->ops_dec_mod[m], s);
      }
    }
  }
  // 3. pack the results of the requested modalities, 4. ONE device -> host copy
  int64_t total = 0;
  for (int m = 0; m < M; ++m) {
    if (!(mask & (1u << m))) continue;
    const Mod& d = c->mods[m];
    if (kind == INF_TRANSFORM) {
      CUDA_OK(cudaMemcpy2DAsync(c->inf.pack + total, (size_t)nz * 4, d.hd, (size_t)d.nh * 4, (size_t)nz * 4, (size_t)B,
                                cudaMemcpyDeviceToDevice, s));
      total += B * nz;
    } else {
      CUDA_OK(cudaMemcpy2DAsync(c->inf.pack + total, (size_t)d.ni * 4, d.xh, (size_t)d.nip * 4, (size_t)d.ni * 4, (size_t)B,
                                cudaMemcpyDeviceToDevice, s));
      total += B * d.ni;
    }
  }
  CUDA_OK(cudaMemcpyAsync(c->inf.pin_out, c->inf.pack, (size_t)total * 4, cudaMemcpyDeviceToHost, s));
}
}  // namespace

extern "C" {

int vaeassoc_infer_host(vaeassoc_handle h, int kind, int modality, const float* const* x_host, const float* z_or_eps_host,
                        float* const* out_host) {
  API_BEGIN(h)
  const int M = h->cfg.n_modalities;
  const int64_t B = h->cfg.batch_size;
  const int nz = h->cfg.n_z;
  if (kind < INF_TRANSFORM || kind > INF_RECONSTRUCT) fail("kind must be 0 (transform), 1 (generate) or 2 (reconstruct)");
  if (modality >= M) fail("modality %d out of range", modality);
  if (!out_host) fail("out_host is null");
  const unsigned mask = modality < 0 ? (1u << M) - 1u : 1u << modality;
  if (kind == INF_GENERATE && !z_or_eps_host) fail("generate needs z_host [batch, n_z]");
  const bool eps_given = kind == INF_RECONSTRUCT && z_or_eps_host != nullptr;
  // inputs -> pinned staging (the graph's H2D nodes read it)
  int64_t off = 0;
  for (int m = 0; m < M; ++m) {
    const int64_t n = B * h->mods[m].ni;
    if (kind != INF_GENERATE && (mask & (1u << m))) {
      if (!x_host || !x_host[m]) fail("x_host[%d] is null", m);
      memcpy(h->inf.pin_in + off, x_host[m], (size_t)n * 4);
    }
    if ((mask & (1u << m)) && !out_host[m]) fail("out_host[%d] is null", m);
    off += n;
  }
  if (z_or_eps_host && kind != INF_TRANSFORM) memcpy(h->inf.pin_in + off, z_or_eps_host, (size_t)(B * nz) * 4);
  cudaStream_t s = h->stream;
  refresh_shadow(h, s);
  if (h->cfg.use_graph) {
    cudaGraphExec_t& g = h->inf.graph[kind][mask][eps_given ? 1 : 0];
    int& nodes = h->inf.nodes[kind][mask][eps_given ? 1 : 0];
    if (!g) nodes = capture(h, &g, [&](cudaStream_t cs) { enqueue_infer(h, kind, mask, eps_given, cs); });
    CUDA_OK(cudaGraphLaunch(g, s));
    h->launches += nodes;
  } else {
    enqueue_infer(h, kind, mask, eps_given, s);
  }
  CUDA_OK(cudaStreamSynchronize(s));
  int64_t total = 0;
  for (int m = 0; m < M; ++m) {
    if (!(mask & (1u << m))) continue;
    const int64_t n = B * (kind == INF_TRANSFORM ? nz : h->mods[m].ni);
    memcpy(out_host[m], h->inf.pin_out + total, (size_t)n * 4);
    total += n;
  }
  API_END(h)
}

}  // extern "C"

// ---- peer-memory data-parallel step (peer_adam.cu) ------------------------------------------------------------------
namespace {
struct PeerBlob {                 // what a rank publishes: VAEASSOC_PEER_BLOB_BYTES
  cudaIpcMemHandle_t handle;      // 64 bytes: the cudaMalloc block that contains the arena
  int64_t offset;                 // arena start inside that block (small arenas are sub-allocated)
  int64_t arena_floats;
  int64_t n_flat;
  int32_t device, pad;
};
static_assert(sizeof(PeerBlob) <= VAEASSOC_PEER_BLOB_BYTES, "peer blob must fit the ABI constant");

void peer_quiesce(Ctx* c) {
  if (!c->peer.on || !c->peer.pending) return;
  enqueue_peer_wait(c, c->stream);
  c->peer.pending = false;
}

void peer_detach(Ctx* c) {
  if (!c->peer.on) return;
  peer_quiesce(c);
  cudaStreamSynchronize(c->stream);
  if (c->peer.tl) {
    unsigned long long h[5] = {0, 0, 0, 0, 0};
    cudaMemcpy(h, c->peer.tl, sizeof h, cudaMemcpyDeviceToHost);
    if (h[0])
      fprintf(stderr, "[peer timeline] rank %d: %llu steps; per step (us): wait for all ranks %.2f | loads + Adam + stores (CTA 0) %.2f | "
              "system fence %.2f | kernel entry -> last CTA out %.2f\n", c->rank, h[0], h[1] * 1e-3 / h[0], h[2] * 1e-3 / h[0],
              h[3] * 1e-3 / h[0], h[4] * 1e-3 / h[0]);
    cudaMemset(c->peer.tl, 0, sizeof h);
  }
  if (c->peer.graph) { cudaGraphExecDestroy(c->peer.graph); c->peer.graph = nullptr; }
  for (int r = 0; r < kMaxPeers; ++r) {
    if (c->peer.mapped[r] && r != c->rank) cudaIpcCloseMemHandle(c->peer.mapped[r]);
    c->peer.base[r] = nullptr; c->peer.mapped[r] = nullptr;
  }
  c->peer.mc = nullptr;
  c->peer.on = false;
}

// moves the flat buffers (and the arrival words) into `new_arena`, memory provided by the caller (symmetric memory that
// every rank of the job has mapped): contents are copied, every schedule is rebuilt around the new addresses
void arena_adopt(Ctx* c, float* new_arena) {
  if (new_arena == c->arena) return;
  if (reinterpret_cast<uintptr_t>(new_arena) & 255) fail("the arena must be 256-byte aligned");
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpy(new_arena, c->arena, (size_t)c->arena_floats * sizeof(float), cudaMemcpyDeviceToDevice));
  const ptrdiff_t delta = reinterpret_cast<char*>(new_arena) - reinterpret_cast<char*>(c->arena);
  char* lo = reinterpret_cast<char*>(c->arena);
  char* hi = lo + (size_t)c->arena_floats * sizeof(float);
  for (auto& g : c->guards)
    if (g.first >= lo && g.first < hi) g.first += delta;        // the guards between the flat buffers travel with them
  c->p += delta / 4; c->g += delta / 4; c->m += delta / 4; c->v += delta / 4; c->p_tf32 += delta / 4;
  c->peer.flags = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(c->peer.flags) + delta);
  c->arena = new_arena;
  c->shadow_dirty = true;
  build_ops(c);
  CUDA_OK(cudaDeviceSynchronize());
}

// m / v of the shards owned by the peers, pulled through the peer mapping (one-sided: the caller's stream is idle and
// the ranks step in lockstep, so no peer is inside an update)
void refresh_peer_slots(Ctx* c) {
  if (!c->peer.on || !c->peer.slots_stale) return;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    int64_t lo, hi;
    peer_shard(c, r, &lo, &hi);
    if (hi <= lo) continue;
    const size_t bytes = (size_t)(hi - lo) * 16;
    for (float* buf : {c->m, c->v})
      CUDA_OK(cudaMemcpy(buf + 4 * lo, c->peer.base[r] + (buf - c->arena) + 4 * lo, bytes, cudaMemcpyDeviceToDevice));
  }
  c->peer.slots_stale = false;
}
}  // namespace

extern "C" {

int vaeassoc_peer_export(vaeassoc_handle h, void* blob) {
  API_BEGIN(h)
  if (!blob) fail("null blob");
  PeerBlob b;
  memset(&b, 0, sizeof b);
  CUDA_OK(cudaIpcGetMemHandle(&b.handle, h->arena));
  // the handle names the whole cudaMalloc block: find the arena's offset inside it
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &qres) != cudaSuccess || !fp)
    fail("cuMemGetAddressRange entry point not available");
  CUdeviceptr base = 0; size_t size = 0;
  if (reinterpret_cast<RangeFn>(fp)(&base, &size, (CUdeviceptr)h->arena) != CUDA_SUCCESS) fail("cuMemGetAddressRange failed");
  b.offset = (int64_t)((CUdeviceptr)h->arena - base);
  b.arena_floats = h->arena_floats; b.n_flat = h->n_flat; b.device = h->device;
  memset(blob, 0, VAEASSOC_PEER_BLOB_BYTES);
  memcpy(blob, &b, sizeof b);
  API_END(h)
}

int vaeassoc_peer_attach(vaeassoc_handle h, const void* all_blobs) {
  API_BEGIN(h)
  if (!h->comm || h->world < 2) fail("vaeassoc_peer_attach needs an initialised communicator with world >= 2");
  if (h->world > kMaxPeers) fail("the peer-memory step serves at most %d ranks (one NVSwitch box)", kMaxPeers);
  if (!all_blobs) fail("null blobs");
  peer_detach(h);
  CUDA_OK(cudaStreamSynchronize(h->stream));
  const char* src = reinterpret_cast<const char*>(all_blobs);
  try {
    for (int r = 0; r < h->world; ++r) {
      PeerBlob b;
      memcpy(&b, src + (size_t)r * VAEASSOC_PEER_BLOB_BYTES, sizeof b);
      if (b.arena_floats != h->arena_floats || b.n_flat != h->n_flat)
        fail("rank %d has a different parameter layout (%lld floats, ours %lld)", r, (long long)b.n_flat, (long long)h->n_flat);
      if (r == h->rank) { h->peer.base[r] = h->arena; continue; }
      int can = 0;
      CUDA_OK(cudaDeviceCanAccessPeer(&can, h->device, b.device));
      if (!can) fail("device %d cannot access device %d through peer memory", h->device, b.device);
      void* mapped = nullptr;
      CUDA_OK(cudaIpcOpenMemHandle(&mapped, b.handle, cudaIpcMemLazyEnablePeerAccess));
      h->peer.mapped[r] = mapped;
      h->peer.base[r] = reinterpret_cast<float*>(reinterpret_cast<char*>(mapped) + b.offset);
    }
  } catch (...) {
    for (int r = 0; r < kMaxPeers; ++r) {
      if (h->peer.mapped[r]) cudaIpcCloseMemHandle(h->peer.mapped[r]);
      h->peer.mapped[r] = nullptr; h->peer.base[r] = nullptr;
    }
    throw;
  }
  // arrival words and epoch start from zero on every rank (the caller barriers between attach and the first step)
  CUDA_OK(cudaMemset(h->peer.flags, 0, (2 * kMaxPeers + 2) * sizeof(uint32_t)));
  CUDA_OK(cudaDeviceSynchronize());
  h->peer.on = true;
  h->peer.pending = false;
  h->peer.slots_stale = false;
  if (getenv("VAEASSOC_PEER_TIMELINE") && !h->peer.tl) h->peer.tl = h->dalloc<unsigned long long>(8);
  destroy_graphs(h);       // the captured launches bake in the task-queue mode
  API_END(h)
}

int64_t vaeassoc_arena_floats(vaeassoc_handle h) { return h ? h->arena_floats : -1; }

int vaeassoc_peer_attach_symmetric(vaeassoc_handle h, void* local_arena, const void* const* rank_arenas, void* multicast_arena) {
  API_BEGIN(h)
  if (!h->comm || h->world < 2) fail("vaeassoc_peer_attach_symmetric needs an initialised communicator with world >= 2");
  if (h->world > kMaxPeers) fail("the peer-memory step serves at most %d ranks (one NVSwitch box)", kMaxPeers);
  if (!local_arena || !rank_arenas) fail("null arena pointers");
  if (rank_arenas[h->rank] != local_arena) fail("rank_arenas[%d] must be this rank's own arena", h->rank);
  peer_detach(h);
  arena_adopt(h, reinterpret_cast<float*>(local_arena));
  for (int r = 0; r < h->world; ++r) {
    if (!rank_arenas[r]) fail("rank_arenas[%d] is null", r);
    h->peer.base[r] = reinterpret_cast<float*>(const_cast<void*>(rank_arenas[r]));
    h->peer.mapped[r] = nullptr;
  }
  h->peer.mc = getenv("VAEASSOC_DP_NO_MULTICAST") ? nullptr : reinterpret_cast<float*>(multicast_arena);
  h->peer.symmetric = true;
  CUDA_OK(cudaMemset(h->peer.flags, 0, (2 * kMaxPeers + 2) * sizeof(uint32_t)));
  CUDA_OK(cudaDeviceSynchronize());
  h->peer.on = true;
  h->peer.pending = false;
  h->peer.slots_stale = false;
  if (getenv("VAEASSOC_PEER_TIMELINE") && !h->peer.tl) h->peer.tl = h->dalloc<unsigned long long>(8);
  destroy_graphs(h);
  API_END(h)
}

int vaeassoc_peer_multicast(vaeassoc_handle h) { return (h && h->peer.on && h->peer.mc) ? 1 : 0; }

int vaeassoc_peer_detach(vaeassoc_handle h) {
  API_BEGIN(h)
  refresh_peer_slots(h);
  peer_detach(h);
  destroy_graphs(h);
  API_END(h)
}

int vaeassoc_peer_active(vaeassoc_handle h) { return (h && h->peer.on) ? 1 : 0; }

int vaeassoc_debug_guard_check(vaeassoc_handle h, int64_t* n_guards, int64_t* n_corrupt) {
  API_BEGIN(h)
  CUDA_OK(cudaDeviceSynchronize());
  int64_t bad = 0;
  std::vector<unsigned char> host;
  for (const auto* vec : {&h->guards, &h->op_guards}) {
    for (const auto& g : *vec) {
      host.resize(g.second);
      CUDA_OK(cudaMemcpy(host.data(), g.first, g.second, cudaMemcpyDeviceToHost));
      for (unsigned char b : host)
        if (b != (unsigned char)Ctx::kGuardByte) { ++bad; break; }
    }
  }
  if (n_guards) *n_guards = (int64_t)(h->guards.size() + h->op_guards.size());
  if (n_corrupt) *n_corrupt = bad;
  API_END(h)
}

}  // extern "C"

extern "C" {

// ---- checkpoint file: "VAEASSOC" | u32 version | u32 n_tensors | i64 step | per tensor { name[48], role[16], i32 ndim,
// i32 shape[4], i64 count, f32 p[count], f32 m[count], f32 v[count] } -- dense logical tensors in the reference's
// tf.Variable creation order (what tf.train.Saver would write: every variable plus its two Adam slots, vae_assoc.py:70)
namespace {
constexpr char kCkptMagic[8] = {'V', 'A', 'E', 'A', 'S', 'S', 'O', 'C'};
struct File {
  FILE* f = nullptr;
  ~File() { if (f) fclose(f); }
};
void get_dense(Ctx* h, const float* base, const vaeassoc_tensor_info& t, std::vector<float>& out) {
  out.resize((size_t)t.rows * t.cols);
  CUDA_OK(cudaMemcpy2D(out.data(), (size_t)t.cols * 4, base + t.offset, (size_t)t.ld * 4, (size_t)t.cols * 4, (size_t)t.rows,
                       cudaMemcpyDeviceToHost));
}
void put_dense(Ctx* h, float* base, const vaeassoc_tensor_info& t, const std::vector<float>& in) {
  CUDA_OK(cudaMemcpy2D(base + t.offset, (size_t)t.ld * 4, in.data(), (size_t)t.cols * 4, (size_t)t.cols * 4, (size_t)t.rows,
                       cudaMemcpyHostToDevice));
}
}  // namespace

int vaeassoc_save(vaeassoc_handle h, const char* path) {
  API_BEGIN(h)
  if (!path || !path[0]) fail("empty checkpoint path");
  CUDA_OK(cudaStreamSynchronize(h->stream));
  refresh_peer_slots(h);
  File fl; fl.f = fopen(path, "wb");
  if (!fl.f) fail("cannot open %s for writing", path);
  const uint32_t version = 1, n = (uint32_t)h->tensors.size();
  int64_t step = 0;
  CUDA_OK(cudaMemcpy(&step, h->step_dev, sizeof step, cudaMemcpyDeviceToHost));
  bool ok = fwrite(kCkptMagic, 1, 8, fl.f) == 8 && fwrite(&version, 4, 1, fl.f) == 1 && fwrite(&n, 4, 1, fl.f) == 1 &&
            fwrite(&step, 8, 1, fl.f) == 1;
  std::vector<float> buf;
  for (const vaeassoc_tensor_info& t : h->tensors) {
    const int64_t count = t.rows * t.cols;
    ok = ok && fwrite(t.name, 1, 48, fl.f) == 48 && fwrite(t.role, 1, 16, fl.f) == 16 && fwrite(&t.ndim, 4, 1, fl.f) == 1 &&
         fwrite(t.shape, 4, 4, fl.f) == 4 && fwrite(&count, 8, 1, fl.f) == 1;
    for (const float* base : {h->p, h->m, h->v}) {
      get_dense(h, base, t, buf);
      ok = ok && fwrite(buf.data(), 4, (size_t)count, fl.f) == (size_t)count;
    }
  }
  if (!ok || fflush(fl.f) != 0) fail("short write to %s", path);
  API_END(h)
}

int vaeassoc_load(vaeassoc_handle h, const char* path) {
  API_BEGIN(h)
  if (!path || !path[0]) fail("empty checkpoint path");
  File fl; fl.f = fopen(path, "rb");
  if (!fl.f) fail("cannot open %s", path);
  char magic[8]; uint32_t version = 0, n = 0; int64_t step = 0;
  if (fread(magic, 1, 8, fl.f) != 8 || memcmp(magic, kCkptMagic, 8) != 0) fail("%s is not a libvaeassoc checkpoint", path);
  if (fread(&version, 4, 1, fl.f) != 1 || fread(&n, 4, 1, fl.f) != 1 || fread(&step, 8, 1, fl.f) != 1 || version != 1)
    fail("%s: unsupported checkpoint version", path);
  if (n != h->tensors.size()) fail("%s holds %u tensors, the model has %zu", path, n, h->tensors.size());
  // read and validate EVERYTHING before the first device write: a mismatch leaves the handle unchanged
  std::vector<std::vector<float>> data(3 * (size_t)n);
  for (uint32_t i = 0; i < n; ++i) {
    const vaeassoc_tensor_info& t = h->tensors[i];
    char name[48], role[16]; int32_t ndim = 0, shape[4]; int64_t count = 0;
    if (fread(name, 1, 48, fl.f) != 48 || fread(role, 1, 16, fl.f) != 16 || fread(&ndim, 4, 1, fl.f) != 1 ||
        fread(shape, 4, 4, fl.f) != 4 || fread(&count, 8, 1, fl.f) != 1)
      fail("%s: truncated tensor table", path);
    name[47] = 0;
    if (strncmp(name, t.name, 48) != 0) fail("%s: tensor %u is '%s', the model expects '%s'", path, i, name, t.name);
    if (ndim != t.ndim || memcmp(shape, t.shape, sizeof shape) != 0 || count != t.rows * t.cols)
      fail("%s: shape of '%s' does not match the model", path, name);
    for (int k = 0; k < 3; ++k) {
      data[3 * i + k].resize((size_t)count);
      if (fread(data[3 * i + k].data(), 4, (size_t)count, fl.f) != (size_t)count) fail("%s: truncated data of '%s'", path, name);
    }
  }
  CUDA_OK(cudaStreamSynchronize(h->stream));
  for (uint32_t i = 0; i < n; ++i) {
    put_dense(h, h->p, h->tensors[i], data[3 * i]);
    put_dense(h, h->m, h->tensors[i], data[3 * i + 1]);
    put_dense(h, h->v, h->tensors[i], data[3 * i + 2]);
  }
  CUDA_OK(cudaMemcpy(h->step_dev, &step, sizeof step, cudaMemcpyHostToDevice));
  h->shadow_dirty = true;
  API_END(h)
}

int64_t vaeassoc_launch_count(vaeassoc_handle h) { return h ? h->launches : -1; }

int vaeassoc_debug_gemm(vaeassoc_handle h, int kind, int use_tc, int M, int N, int K, const float* A, int64_t lda,
                        const float* B, int64_t ldb, float* C, int64_t ldc, const float* bias, float* bias_grad,
                        const float* aux, int64_t ldaux, int act, int round_out) {
  API_BEGIN(h)
  GemmArgs a;
  a.M = M; a.N = N; a.K = K; a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc; a.bias = bias;
  a.bias_grad = bias_grad; a.aux = aux; a.ldaux = ldaux; a.act = act; a.round_out = round_out;
  if (use_tc == 1) {
    if (!tc_supported(kind, a)) fail("shape not served by the tcgen05 path");
    // a throw-away one-problem plan through the same persistent tile kernel the train step uses
    char err[256] = {0};
    GroupPlan* plan = group_create();
    const int dsite = group_begin(plan);
    const int prob = group_add_problem(plan, kind, a, err, sizeof err);
    if (prob < 0) { group_destroy(plan); fail("tcgen05 plan failed: %s", err); }
    const int tm = group_problem_tiles_m(plan, prob), tn = group_problem_tiles_n(plan, prob);
    const int kb = group_problem_kblocks(plan, prob);
    const int kpr = group_problem_kb_per_rowblock(plan, prob);
    const int rbs = (kb + kpr - 1) / kpr;
    const int splits = kind == KIND_TN ? std::max(1, std::min(rbs, (kNumSMs / 2) / std::max(1, tm * tn))) : 1;
    const int per = (rbs + splits - 1) / splits;
    for (int r0 = 0; r0 < (kind == KIND_TN ? rbs : 1); r0 += per)
      for (int i = 0; i < tm; ++i)
        for (int j = 0; j < tn; ++j)
          group_add_task(plan, prob, i, j, kind == KIND_TN ? r0 * kpr : 0,
                         kind == KIND_TN ? std::min(per * kpr, kb - r0 * kpr) : kb, -1, 0, 0, -1, 0, -1);
    group_set_counters(plan, h->gsync, h->n_ctr);
    if (!group_end(plan, err, sizeof err) || !group_upload(plan, err, sizeof err)) { group_destroy(plan); fail("%s", err); }
    group_launch(plan, dsite, h->gsync + h->n_ctr + 2 * (h->max_sites - 1), 0, 0, dyn_first(h), h->stream);
    float* cws = nullptr;
    if (kind == KIND_TN && bias_grad) {
      const int64_t nws = colsum_ws_floats(K, N);
      CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&cws), (size_t)nws * 4));
      CUDA_OK(cudaMemsetAsync(cws, 0, (size_t)nws * 4, h->stream));
      launch_colsum(B, ldb, K, N, bias_grad, cws, h->stream);
    }
    CUDA_OK(cudaStreamSynchronize(h->stream));
    if (cws) cudaFree(cws);
    if (getenv("VAEASSOC_TC_TIMELINE"))
      group_debug_timeline(plan, dsite, h->gsync + h->n_ctr + 2 * (h->max_sites - 1), 0, 0, h->stream);
    group_destroy(plan);
  } else {
    const bool sk = use_tc == 0 && skinny_supported(kind, a);
    const int64_t nws = sk ? gemm_skinny_ws_floats(kind, a) : (kind == KIND_TN ? gemm_tn_simt_ws_floats(a) : 0);
    if (nws > 0) CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&a.ws), (size_t)nws * 4));
    if (sk) {
      launch_gemm_skinny(kind, a, h->stream);
    } else {
      switch (kind) {
        case KIND_NN: launch_gemm_nn_simt(a, h->stream); break;
        case KIND_NT: launch_gemm_nt_simt(a, h->stream); break;
        default: launch_gemm_tn_simt(a, h->stream); break;
      }
    }
    CUDA_OK(cudaStreamSynchronize(h->stream));
    if (a.ws) cudaFree(a.ws);
  }
  CUDA_OK(cudaGetLastError());
  API_END(h)
}

int vaeassoc_profile_step(vaeassoc_handle h, const float* const* x_dev, const int64_t* ld, const float* eps_dev,
                          char* names, float* ms, double* flops, double* bytes, int capacity) {
  if (!h) return -1;
  try {
    check_handle(h);
    std::lock_guard<std::mutex> lock(h->mu);
    if (!x_dev) fail("x_dev is null");
    cudaStream_t s = h->stream;
    stage_inputs(h, x_dev, ld, eps_dev, s);
    refresh_shadow(h, s);
    h->g_zero = false;
    std::vector<Op> all;
    {
      Op z; z.name = "zero_grads"; z.bytes = 4.0 * (h->n_flat + 32);
      Ctx* c = h;
      z.run = [c](cudaStream_t st) { CUDA_OK(cudaMemsetAsync(c->g, 0, (size_t)(c->n_flat + 32) * sizeof(float), st)); };
      all.push_back(z);
    }
    h->lat_mode_elt = elt_mode(h);
    if (one_mode(h)) {
      Ctx* c = h;
      Op o; o.name = "seg_step";
      for (auto* v : {&h->ops_fwd_enc, &h->ops_fwd_dec, &h->ops_bwd_dec, &h->ops_bwd_enc}) for (auto& op : *v) o.flops += op.flops;
      o.run = [c](cudaStream_t st) { launch_seg(c, c->seg_step, st, 0); c->launches -= 1; };
      all.push_back(o);
      h->recon_stale = (1u << h->cfg.n_modalities) - 1u;
    } else if (elt_mode(h)) {
      Ctx* c = h;
      auto seg_op = [&](const char* name, const Ctx::Seg* sg, std::initializer_list<std::vector<Op>*> members) {
        Op o; o.name = name;
        for (auto* v : members) for (auto& op : *v) o.flops += op.flops;
        o.run = [c, sg](cudaStream_t st) { launch_seg(c, *sg, st); c->launches -= 1; };
        all.push_back(o);
      };
      seg_op("seg_fwd", &h->seg_fwd, {&h->ops_fwd_enc, &h->ops_fwd_dec});
      for (auto& op : h->ops_loss) all.push_back(op);
      for (auto& op : h->ops_colsum_dec) all.push_back(op);
      seg_op("seg_bwd", &h->seg_bwd, {&h->ops_bwd_dec, &h->ops_bwd_enc});
    } else if (h->fused) {
      // the train step's own launches: four fused segments + the elementwise kernels between them
      Ctx* c = h;
      auto seg_op = [&](const char* name, const Ctx::Seg* sg, std::initializer_list<std::vector<Op>*> members) {
        Op o; o.name = name;
        for (auto* v : members) for (auto& op : *v) o.flops += op.flops;
        o.run = [c, sg](cudaStream_t st) { launch_seg(c, *sg, st); c->launches -= 1; };
        all.push_back(o);
      };
      seg_op("seg_fwd_enc", &h->seg_enc, {&h->ops_fwd_enc});
      for (auto& op : h->ops_latent_fwd) all.push_back(op);
      seg_op("seg_fwd_dec", &h->seg_dec, {&h->ops_fwd_dec});
      for (auto& op : h->ops_loss) all.push_back(op);
      for (auto& op : h->ops_colsum_dec) all.push_back(op);
      seg_op("seg_bwd_dec", &h->seg_bwd_dec, {&h->ops_bwd_dec});
      for (auto& op : h->ops_latent_bwd) all.push_back(op);
      for (auto& op : h->ops_colsum_enc) all.push_back(op);
      seg_op("seg_bwd_enc", &h->seg_bwd_enc, {&h->ops_bwd_enc});
    } else {
      for (auto* v : {&h->ops_fwd_enc, &h->ops_latent_fwd, &h->ops_fwd_dec, &h->ops_loss, &h->ops_bwd_dec, &h->ops_latent_bwd, &h->ops_bwd_enc})
        for (auto& op : *v) all.push_back(op);
    }
    {
      Ctx* c = h;
      Op f; f.name = "finalize";
      f.run = [c](cudaStream_t st) {
        FinalizeArgs fa = one_mode(c) ? c->fin_one : finalize_args(c, 1);
        fa.advance = 1;
        launch_finalize(fa, st);
      };
      all.push_back(f);
      Op a; a.name = "adam"; a.bytes = (c->cfg.precision == VAEASSOC_TF32 ? 32.0 : 28.0) * c->n_flat;
      a.run = [c](cudaStream_t st) { launch_adam(adam_args(c), st); };
      all.push_back(a);
    }
    if (h->fused && getenv("VAEASSOC_TC_TIMELINE")) {
      CUDA_OK(cudaStreamSynchronize(s));
      std::vector<const Ctx::Seg*> segs = {&h->seg_enc, &h->seg_dec, &h->seg_bwd_dec, &h->seg_bwd_enc};
      if (elt_mode(h)) segs = {&h->seg_fwd, &h->seg_bwd};
      if (one_mode(h)) segs = {&h->seg_step};
      for (const Ctx::Seg* sg : segs)
        group_debug_timeline(h->gplan, sg->site, h->gsync + h->n_ctr + 2 * sg->site, sg->reset_first, sg->reset_count, s);
    }
    const int n = (int)std::min<size_t>(all.size(), (size_t)capacity);
    std::vector<cudaEvent_t> ev(all.size() + 1);
    for (auto& e : ev) CUDA_OK(cudaEventCreate(&e));
    CUDA_OK(cudaEventRecord(ev[0], s));
    // each op runs kReps times back to back between its two events (an event pair around ONE launch measures
    // mostly the launch gap: ~8 us against kernels of 5-30 us); accumulating ops (wgrad) just accumulate kReps times
    constexpr int kReps = 4;
    for (size_t i = 0; i < all.size(); ++i) {
      for (int r = 0; r < kReps; ++r) all[i].run(s);
      h->launches += kReps * all[i].launches;
      CUDA_OK(cudaEventRecord(ev[i + 1], s));
    }
    CUDA_OK(cudaStreamSynchronize(s));
    for (int i = 0; i < n; ++i) {
      float t = 0.f;
      CUDA_OK(cudaEventElapsedTime(&t, ev[i], ev[i + 1]));
      ms[i] = t / kReps;
      if (flops) flops[i] = all[i].flops;
      if (bytes) bytes[i] = all[i].bytes;
      snprintf(names + (size_t)i * 32, 32, "%s", all[i].name.c_str());
    }
    for (auto& e : ev) cudaEventDestroy(e);
    return n;
  } catch (const std::exception& e) {
    h->err = e.what();
    return -1;
  }
}

}  // extern "C"
