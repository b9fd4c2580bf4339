// Persistent, grouped, dataflow tcgen05 / TMA GEMM for sm_100a -- the tensor-core path of the assoc-VAE train step.
//
// One launch executes a LIST OF TILE TASKS drawn from several dense-layer contractions ("problems") of the step:
//
//   kind NN  C[M,N]  = act(A[M,K] . B[K,N] + bias)          forward   (vae_assoc.py:187-188,203-204,282-283,295-303)
//   kind NT  C[M,N]  = (A[M,K] . B[N,K]^T) (*) act'(aux)    dgrad     (autodiff of the above, :373-374)
//   kind TN  C[M,N] += A[K,M]^T . B[K,N]                    wgrad     (K = batch, split into tasks, TMA reduce-add)
//
//   elementwise tasks (no contraction): the latent stage of a row block and its backward (latent.cuh), the cost finalize
//
// Why: at the reference sizes one layer is 0.5-6 GFLOP, i.e. 3-8 us of tensor time; launched one kernel per layer
// (34 GEMMs + 12 column sums per step) the fixed cost of a launch -- grid start, barrier / TMEM set-up, first
// operand latency, epilogue drain, ~9 us -- exceeds the useful work.  Rows of the batch are independent through the
// whole network, so the WHOLE GRADIENT STEP of both modalities (round 1: each of its four segments) is ONE launch:
// each task names the row-block counters it must see complete before its operands may be loaded (the tiles of the
// producing layer over the same 256 rows) and the counter that is bumped when its own output is globally visible.
// Tasks are listed in dependency order and handed out IN THAT ORDER from a global atomic queue to whichever cluster
// is free, so the smallest unfinished task is always held by a running cluster and is never blocked -- no deadlock,
// whatever share of the SMs other streams (NCCL) occupy.
//
// Tile engine (per cluster of two CTAs, `tcgen05.mma.cta_group::2`, UMMA 256 x BN x 8, kind::tf32), 11 warps per CTA:
//   warp 0     TMA producer: waits the task's counters, then streams A (its 128 rows) and B (its BN/2 columns)
//              k-blocks of 32 into a ring of 4 stages (5 / 6 when every tile of the launch is narrow; 3 stages of 64-deep
//              k-blocks at small batches, see `ring_class`); both CTAs' loads complete on the leader's `full` barrier.
//              A consumer of a half-tile hand-over (TF_HALF) streams the k-blocks of every producing tile's first half
//              of chunks as soon as those are published, the rest after the full counters.
//              The leader's producer is also the scheduler: it pops the queue one task ahead and publishes each task
//              index through an 8-slot ring (shared memory of both CTAs, mbarrier full / empty) to every role
//   warp 1     leader CTA: MMA issuer; accumulators double-buffered in TMEM (2 x 256 columns) so the epilogue of
//              task i overlaps the main loop of task i+1; `tcgen05.commit` multicasts to both CTAs
//   warps 2-9  epilogue: tcgen05.ld (thread = row, 32 columns) -> bias / activation / relu' from 1-bit masks / act'(aux
//              tile via TMA, three boxes in flight per warp, processed in place) / reconstruction loss + d cost / d a
//              against the target tile (TF_LOSS: the decoders' output layer) -> swizzled smem box -> coalesced 128-bit
//              global stores (NN, NT: 8 lanes per 128-byte row segment) or TMA reduce-add (TN, split-K NN / NT); then
//              release the accumulator and ARRIVE ON A CTA-LOCAL BARRIER (half-tile and full).  They also execute the
//              elementwise tasks.  Nothing in the chunk loop waits on a fresh global / TMA round trip: bias values,
//              mask words and aux boxes are requested before the accumulator is complete
//   warp 10    signal warp: waits for those barriers, fences once (fence.acq_rel.gpu, cumulative over the epilogue
//              warps' stores) and adds the CTA's eight arrivals to the row-block counter -- the MEMBAR.GPU of a release
//              (~0.9 us) is off the epilogue warps
// The kernel allocates 168 registers per thread (352 threads allocate like 384) and the chunk loop of the epilogue sits
// at that limit: every variant that spilled ~150 bytes inside the loop lost 15-20 % on every configuration.
// Operands are fp32 in HBM, rounded to tf32 by their producers.  TMA zero-fills loads past M / N / K and clips
// stores (in 16-byte units: columns N..roundup4(N) receive zeros), so only 16-byte row pitches are required.
//
// Shared-memory operand layouts (canonical UMMA layouts):
//   K-major  operand: tile [R rows][32 k] -> R rows of 128 B, one TMA box {32, R}; SWIZZLE_128B, SBO = 1024 B;
//            the four K=8 MMAs of a stage advance the start address by 32 B.
//   MN-major operand: tile [R/32 chunks][32 k rows][32 mn]; for 32-bit operands the only UMMA layout is
//            SWIZZLE_128B_BASE32B (TMA mode SWIZZLE_128B_ATOM_32B): LBO = 4096 B, SBO = 512 B; the four MMAs of a
//            stage advance the start by 1024 B.  ONE TMA instruction fills it: the matrix is described as a 3-D
//            tensor {32 mn, k, mn-chunk} with strides {4 B, pitch, 128 B} and the box is {32, 32, R/32}.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "latent.cuh"

namespace vaeassoc {

namespace {

constexpr int BM_CTA = 128;      // rows of the tile per CTA; UMMA M = 256 over the pair
constexpr int BM = 2 * BM_CTA;
constexpr int BK = 32;           // fp32 elements per stage = one 128-byte swizzle row
constexpr int UMMA_K = 8;        // tf32: 32 bytes per instruction
constexpr int kEpiWarps = VAEASSOC_EPI_WARPS;   // kSlots warps per TMEM lane quarter; slot s takes chunks s, s + kSlots, ...
constexpr int kSlots = kEpiWarps / 4;
static_assert(kEpiWarps % 4 == 0 && kSlots >= 1 && kSlots <= 4, "epilogue warps come in groups of four (one per TMEM lane quarter)");
constexpr int kThreads = 64 + 32 * kEpiWarps + 32;   // producer, MMA issuer, epilogue warps, signal warp
constexpr int kSignalWarp = 2 + kEpiWarps;
constexpr int kDoneRing = 4;                          // tile-completion barriers between the epilogue warps and the signal warp
constexpr int kBarRegion = 640;                       // bytes: barriers, TMEM slot, task ring, signal-warp sequence word
constexpr int kStages = 4;        // operand ring at the widest tile (32 KB per stage)
constexpr int kStagesMax = 8;     // launches whose tiles are all narrow cut the same 128 KB into more, smaller stages
constexpr int A_BYTES = BM_CTA * BK * 4;          // 16 KB
constexpr int B_BYTES_MAX = 128 * BK * 4;         // BN/2 <= 128 columns
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;
constexpr int CHUNK_BYTES = 32 * 32 * 4;          // one 32 x 32 fp32 epilogue box
constexpr int kEpiBufs = kEpiWarps <= 8 ? 3 : 1;  // rotating boxes per epilogue warp: aux tile in, result out (in place)
constexpr int EPI_WARP_BYTES = kEpiBufs * CHUNK_BYTES;
constexpr int kAccCols = 256;                     // TMEM columns per accumulator
constexpr int kTmemCols = 2 * kAccCols;
#ifndef VAEASSOC_TIMELINE
#define VAEASSOC_TIMELINE 0            // per-task %globaltimer stamps (VAEASSOC_TC_TIMELINE): `python -m vae_assoc_b200.build --timeline`
#endif
constexpr bool kTimeline = VAEASSOC_TIMELINE != 0;
#ifndef VAEASSOC_EPI_DEBUG
#define VAEASSOC_EPI_DEBUG 0
#endif
constexpr bool kEpiDebug = VAEASSOC_EPI_DEBUG != 0;   // (compiled out by default: the stamps cost registers in every role)
constexpr int kTL = 16;                          // debug timeline: 64-bit stamps per task
constexpr int kSched = 8;                        // depth of the task-index ring
constexpr int kBiasStrip = 128;                  // bytes per epilogue warp: the 32 bias values of the current chunk
constexpr int SMEM_BYTES = kStages * STAGE_BYTES + kEpiWarps * EPI_WARP_BYTES + kBarRegion /*barriers, task ring*/ +
                           kEpiWarps * kBiasStrip + 1024 /*align*/;
static_assert(SMEM_BYTES <= 232448, "more than the 227 KB a CTA may opt into");

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on a barrier of any CTA of the cluster; release: orders this thread's earlier shared-memory writes
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
// no memory ordering (a release at cluster scope costs a MEMBAR.GPU + L1 invalidate); `dep` is a register the arrive
// must wait for (e.g. the value just read from the slot this arrive hands back)
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar_cluster, uint32_t dep) {
  asm volatile("{\n\t.reg .b32 dummy;\n\tmov.b32 dummy, %1;\n\t"
               "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];\n\t}" ::"r"(bar_cluster), "r"(dep) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// acquire at cluster scope: the waiter reads shared memory written by the peer CTA before its arrive
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded waits: a protocol bug must surface as a launch failure, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    if (spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait_cluster(bar, parity); ++spins) {
    if (spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_complete() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// asynchronous form: the registers are valid only after tmem_ld_wait(v); the wait names them as in/out operands so
// that the compiler keeps every use behind it
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                 "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// publishes one word into the shared memory of a CTA of the cluster and completes 4 transaction bytes on that CTA's
// barrier -- the data is visible to whoever observes the barrier phase (async proxy, like a TMA load): no fence needed
__device__ __forceinline__ void st_async_u32(uint32_t addr_cluster, uint32_t v, uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];"
               ::"r"(addr_cluster), "r"(v), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster_relaxed(uint32_t bar_cluster, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_release_gpu_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout): start [0,14) >>4, LBO [16,30) >>4,
// SBO [32,46) >>4, version [46,48) = 1, layout_type [61,64): 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_k_major(uint32_t addr) { return make_desc(addr, 16, 1024, 2); }
__device__ __forceinline__ uint64_t desc_mn_major(uint32_t addr, uint32_t k_rows = BK) { return make_desc(addr, k_rows * 128, 512, 1); }

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format [4,6)=1 (F32), a/b_format [7,10),[10,13)=2 (TF32),
// a_major bit 15, b_major bit 16 (1 = MN-major), n_dim [17,23) = N>>3, m_dim [24,29) = M>>4
__device__ __forceinline__ uint32_t make_idesc(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// byte offset of 16-byte chunk j of row r inside a 32 x 32 fp32 box written / read by TMA with SWIZZLE_128B
__device__ __forceinline__ uint32_t swz(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

}  // namespace

// one dense-layer contraction; lives in global memory (the TMA unit reads the tensor maps from there)
struct alignas(64) GProblem {
  CUtensorMap map_a, map_b, map_c, map_aux;
  int M, N, K, BN;           // output rows, output cols, contraction length, tile width (64, 128, 192 or 256)
  int a_mn, b_mn;            // operand is MN-major (M / N index contiguous in HBM)
  int reduce;                // epilogue adds into C (TMA reduce-add): wgrad tasks share an output tile
  int act, round_out, has_aux;
  int loss_binary;           // loss-fused output layer: Bernoulli cross-entropy (else Gaussian l2)
  float loss_scale;
  int k64;                   // operand maps deliver 64-deep k-blocks (two 32-deep chunks per TMA instruction), see kDeepK
  const float* bias;
  float* colsum;             // epilogue adds the column sums of its output tile here (bias gradient), or null
  float* c_ptr;              // output matrix for the direct (non-TMA) stores of the NN / NT epilogue
  long long ldc;             // its row pitch in floats (multiple of 4)
  uint32_t* mask_out;        // relu forward: bit (row, col) = output > 0, 32 columns per word, or null
  const uint32_t* mask_in;   // relu dgrad: the same words replace the aux tile, or null
  long long ldmask;          // words per mask row
  float* loss_partials;      // loss-fused output layer (aux tile = target x, C = d a): per-warp loss sums, or null
};
static_assert(sizeof(GProblem) % 64 == 0, "tensor maps must stay 64-byte aligned inside the array");

// one output tile (256 x BN) of one problem over k-blocks [kb0, kb0 + nkb); 64 bytes, the problem's scalars are
// repeated here so that a role needs ONE load per task (prefetched while the previous task runs)
struct alignas(16) GTask {
  int problem, m_blk, n_blk, kb0;
  int nkb;
  int wait_ctr, wait_cnt, wait_val;   // operands are ready once counters[wait_ctr .. +wait_cnt) have all reached wait_val
  int wait2_ctr, wait2_val;           // a second counter (or -1)
  int signal_ctr;                     // bumped by each of the 16 epilogue warps when the tile is globally visible, or -1
  int bn;                             // tile width of the problem
  int flags;                          // bit 0 a_mn, 1 b_mn, 2 reduce, 3 has_aux, 4 round_out, 5 colsum, 6 mask_out, 7 mask_in
  int act, M, N;
};
static_assert(sizeof(GTask) == 64, "GTask is loaded as four 16-byte words");
enum { TF_A_MN = 1, TF_B_MN = 2, TF_REDUCE = 4, TF_AUX = 8, TF_ROUND = 16, TF_COLSUM = 32, TF_MASK_OUT = 64, TF_MASK_IN = 128,
       // elementwise tasks (no contraction: nkb = 0; the epilogue warps of the pair execute them over the 256 rows of
       // row block m_blk; the producer and the MMA issuer skip them): latent forward / backward (latent.cuh)
       TF_ELT_LATENT_FWD = 256, TF_ELT_LATENT_BWD = 512,
       // NN task of a decoder's output layer with the reconstruction loss fused into its epilogue (aux tile = target)
       TF_LOSS = 1024,
       // elementwise task: cost finalize (one warp sums the block partials of the loss / latent tasks; kb0 = number of
       // counters behind wait2_ctr)
       TF_ELT_FINALIZE = 2048, TF_ELT = TF_ELT_LATENT_FWD | TF_ELT_LATENT_BWD | TF_ELT_FINALIZE,
       // half-tile hand-over between dependent row-wise layers: a producing tile publishes its first half of 32-column
       // chunks (counter + half_off) before its second half is out of TMEM; the consuming tile streams the k-blocks of
       // every producing tile's first half, then waits for the full counters and streams the rest (k order is free in a
       // contraction).  TF_HALF: consumer (wait2_ctr = half counter, wait2_val = k-blocks per producing tile);
       // TF_SIG_HALF: producer
       TF_HALF = kTaskHalf, TF_SIG_HALF = kTaskSigHalf };
// Rows of a row-wise (NN / NT) tile per CTA of the pair.  A batch of at most 256 rows is ONE row block: instead of 128 rows
// in the leader CTA and the rest (at B = 100: none) in its peer, each CTA takes half of them (rounded up to 8), so that
// either CTA's TMA stream carries half of the A operand -- at small batches a main loop is bound by the bytes one SM
// can pull per clock (~48), and the zero-filled rows of a 128-row box count like real ones.  The peer's rows then start
// at rows_per_cta(M) instead of 128; TMEM lane = local row; rows at or past rows_per_cta within a CTA are dead.
__host__ __device__ __forceinline__ int rows_per_cta(int M, bool small_rows) {
  return (small_rows && M <= 256) ? max(8, (((M + 1) >> 1) + 7) & ~7) : 128;
}
// first-half chunk count of a tile with `n` chunks (both epilogue slots take the same number of first-half chunks)
__host__ __device__ __forceinline__ int half_chunks(int n) { const int h = (((n + 1) >> 1) + 1) & ~1; return h < n ? h : n; }
// (half_chunks() assumes two epilogue slots per TMEM lane quarter: the host only plans hand-overs when kGroupHalfOk)


namespace {

// polls with RELAXED loads: an acquire load is LDG.STRONG + CCTL.IVALL (invalidate the whole L1) per iteration, and a
// producer lane spinning on it every ~100 ns slowed every shared-memory / shuffle / store instruction of the epilogue
// warps on the same SM by 5x (measured: 1.4 us per 32x32 chunk instead of 0.27).  The caller issues ONE
// fence.acq_rel.gpu after its last counter has been observed.
__device__ __forceinline__ void wait_counter(const uint32_t* ctr, uint32_t need) {
  const long long t0 = clock64();
  while (ld_relaxed_gpu(ctr) < need) {
    __nanosleep(100);
    if (clock64() - t0 > 4000000000ll) __trap();    // ~2 s: a scheduling bug must not hang the GPU
  }
}

// one lane of the (converged) warp.  With `elect.sync` the compiler knows the guarded region runs on a single thread
// and issues UTMALDG / UTMASTG / UTCHMMA straight from uniform registers; behind `if (lane == 0)` every such
// instruction whose operands came from memory (task fields) was wrapped in an R2UR.BROADCAST waterfall loop
// (~130 cycles per TMA issue: the producer needed 0.64 us per k-block instead of 0.4)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// value of lane 0 in every lane
__device__ __forceinline__ int bcast(int v) { return __shfl_sync(0xffffffffu, v, 0); }

// one 32-column chunk of a loss-fused output tile, this lane's row: v[j] = pre-activation bits in, d cost / d a bits out
// (rounded to tf32 when `round`); x = the target row in the swizzled aux box (16-byte column j at xrow + ((j ^ xsw) << 4));
// returns the row's loss over the chunk.  Columns >= ncols (past N, or the whole row past M) contribute nothing and
// leave exact zeros (the column sums of the bias gradient read the staged tile).
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <bool BINARY, bool FULL>
__device__ __forceinline__ float loss_chunk(uint32_t (&v)[32], uint32_t xrow, uint32_t xsw, float scale, bool round, int ncols) {
  float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 xv = lds128(xrow + (((uint32_t)j ^ xsw) << 4));
    const float xs4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float a = __uint_as_float(v[4 * j + k]), x = xs4[k];
      float da, l;
      if (BINARY) {
        // x_hat = sigmoid(a); loss = -(x log(1e-3 + x_hat) + (1 - x) log(1e-3 + 1 - x_hat))   vae_assoc.py:321-324
        // (in log2 units here, scaled once per task); d loss / d a over ONE reciprocal.  (A four-MUFU form -- one
        // reciprocal of t P Q with t = 1 + e^-a, P = eps t + 1, Q = (1 + eps) t - 1 giving both x_hat and the gradient
        // factor -- measured no faster: 0.2764 against 0.2743 ms per step, three more live registers.)
        const float xh = rcp_approx(1.0f + ex2_approx(a * -1.4426950408889634f));
        const float pp = kCeEps + xh;
        const float qq = (kCeEps + 1.0f) - xh;            // evaluation order of :323
        const float omx = 1.0f - x;
        l = x * lg2_approx(pp) + omx * lg2_approx(qq);
        da = (scale * (omx * pp - x * qq)) * ((xh * (1.0f - xh)) * rcp_approx(pp * qq));
      } else {
        const float d = a - x;                            // tf.nn.l2_loss(x_hat - x), x_hat = a   :327-328
        da = scale * d;
        l = d * d;
      }
      uint32_t bits = __float_as_uint(da);
      if (round) bits = (bits + 0x1000u) & 0xffffe000u;
      if (FULL) {
        v[4 * j + k] = bits;
      } else {
        const bool live = 4 * j + k < ncols;
        l = live ? l : 0.0f;
        v[4 * j + k] = live ? bits : 0u;
      }
      if (k & 1) acc1 += l; else acc0 += l;
    }
  }
  return (BINARY ? -0.6931471805599453f : 0.5f) * (acc0 + acc1);
}

__device__ __forceinline__ GTask load_task(const GTask* __restrict__ tasks, int t) {
  GTask tk;
  if (t >= 0) {
    const int4* src = reinterpret_cast<const int4*>(tasks + t);
    int4* dst = reinterpret_cast<int4*>(&tk);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = __ldg(src + i);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) reinterpret_cast<int*>(&tk)[i] = 0;
  }
  return tk;
}

// the problems of one launch travel as a __grid_constant__ kernel parameter: the TMA unit fetches tensor maps from
// param / const space through its descriptor cache; with the maps in plain global memory every cp.async.bulk.tensor
// paid a descriptor fetch (~250 cycles, measured: 1 340 instead of 680 cycles per k-block)
template <int NP>
struct GParams { GProblem p[NP]; };

template <int NP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_group_kernel(const __grid_constant__ GParams<NP> params, const __grid_constant__ GElem elem,
                  const GTask* __restrict__ tasks, int ntasks,
                  uint32_t* __restrict__ counters, uint32_t* __restrict__ queue, int reset_first, int reset_count,
                  int mode, unsigned long long* __restrict__ tl, int half_off) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = base + kStages * STAGE_BYTES;
  const uint32_t bar_base = epi_base + kEpiWarps * EPI_WARP_BYTES;
  const uint32_t strip_base = bar_base + (uint32_t)kBarRegion;
  const int dynamic_first = mode & 1;
  // epilogue variants behind environment switches (bias by 32 shuffles, TMA stores, timing experiments that drop parts of
  // the epilogue) exist only in the debug build (-DVAEASSOC_EPI_DEBUG=1): each costs a live register or predicate in the
  // chunk loop, whose allocation sits at the 168-register limit
  const bool bias_smem = !kEpiDebug || (mode & 2) != 0, tma_store = kEpiDebug && (mode & 4) != 0;
  const bool dbg_skip_math = kEpiDebug && (mode & 8) != 0, dbg_skip_store = kEpiDebug && (mode & 16) != 0;
  const int fin_advance = (mode >> 5) & 1;               // the finalize task bumps the Adam step counter
  // operand ring geometry of this launch (host: launch_site): the 128 KB ring holds 4 stages at the widest tile, 5 when
  // no tile of the launch is wider than 128, 6 when none is wider than 64 -- at small batches (one row block, 64-wide
  // tiles) a main loop is bound by the TMA round trip per stage, i.e. by the number of stages in flight
  const bool small_rows = ((mode >> 24) & 1) != 0;      // rows_per_cta(): the plan's tensor maps were built with it
  const int ring_class = (mode >> 6) & 3;
  // ring_class 3 = 64-deep k-blocks (narrow launches at small batches): a stage is [A chunk 0 | A chunk 1] (2 x 16 KB) +
  // [B chunk 0 | B chunk 1] (2 x 4 KB), three stages; ONE TMA instruction per operand fills both chunks.  At small batches
  // a main loop is bound by the producer's instruction rate (~0.23 us per k-block of two TMA instructions, whatever the
  // bytes and the ring depth), so half the instructions per k is half the main loop.
  const bool k64 = ring_class == 3;
  const int nstages = k64 ? 3 : ring_class == 2 ? 6 : ring_class == 1 ? 5 : kStages;
  const uint32_t stage_bytes = k64 ? (uint32_t)(2 * A_BYTES + 2 * 32 * BK * 4) : ring_class == 2 ? (uint32_t)(A_BYTES + 32 * BK * 4) : ring_class == 1 ? (uint32_t)(A_BYTES + 64 * BK * 4) : (uint32_t)STAGE_BYTES;
  const uint32_t b_off = k64 ? 2u * A_BYTES : (uint32_t)A_BYTES;      // B tile of a stage behind its A tile(s)
  const uint32_t stagger_ns = kEpiDebug ? ((uint32_t)mode >> 8) & 0xfffu : 0u;      // VAEASSOC_EPI_STAGGER_NS: the odd chunk slots start this much later
  // barriers: full[s] (leader CTA only), empty[s], tmem_full[2], tmem_empty[2] (leader CTA only), aux[epilogue warp],
  // sched_full[kSched], sched_empty[kSched] (leader CTA only); then the TMEM slot and the task-index ring
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStagesMax + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * kStagesMax + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * kStagesMax + 2 + a); };
  auto aux_bar = [&](int e, int b) { return bar_base + 8u * (2 * kStagesMax + 4 + kEpiBufs * e + b); };
  constexpr int kAuxBars = kEpiBufs * kEpiWarps;
  auto sched_full_bar = [&](int r) { return bar_base + 8u * (2 * kStagesMax + 4 + kAuxBars + r); };
  auto sched_empty_bar = [&](int r) { return bar_base + 8u * (2 * kStagesMax + 4 + kAuxBars + kSched + r); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStagesMax + 4 + kAuxBars + 2 * kSched);
  auto sched_task = [&](int r) { return tmem_slot + 8u + 4u * r; };
  // done[i]: the kEpiWarps epilogue warps of THIS CTA have stored their part of the i-th (mod kDoneRing) signalling tile
  auto done_bar = [&](uint32_t i) { return tmem_slot + 8u + 4u * kSched + 8u * (i % kDoneRing); };
  auto half_bar = [&](uint32_t i) { return tmem_slot + 8u + 4u * kSched + 8u * kDoneRing + 8u * (i % kDoneRing); };
  const uint32_t sig_seq_addr = tmem_slot + 8u + 4u * kSched + 8u * 2 * kDoneRing;   // signalling tiles published so far
  static_assert(8 * (2 * kStagesMax + 4 + kEpiBufs * kEpiWarps + 2 * kSched) + 8 + 4 * kSched + 16 * kDoneRing + 8 <= kBarRegion,
                "barrier region too small");
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // consumers of a published task index: leader MMA lane + peer producer lane + 2 x kEpiWarps epilogue lanes
  constexpr uint32_t kSchedConsumers = 2 + 2 * kEpiWarps + 2;     // ... + the signal warp of either CTA

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStagesMax; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), 2 * kEpiWarps); }
    for (int e = 0; e < kEpiWarps; ++e)
      for (int b = 0; b < kEpiBufs; ++b) mbar_init(aux_bar(e, b), 1);
    for (int r = 0; r < kSched; ++r) { mbar_init(sched_full_bar(r), 1); mbar_init(sched_empty_bar(r), kSchedConsumers); }
    for (int i = 0; i < kDoneRing; ++i) { mbar_init(done_bar(i), kEpiWarps); mbar_init(half_bar(i), kEpiWarps); }
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sig_seq_addr), "r"(0u) : "memory");
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const unsigned long long t_entry = (kTimeline && tl) ? gtimer() : 0ull;

  // consumer side of the task ring: next task index (or -1 when the queue is drained); one calling lane per role
  uint32_t fetches = 0;
  auto next_task = [&]() -> int {
    const int r = (int)(fetches % kSched);
    mbar_wait(sched_full_bar(r), (fetches / kSched) & 1);
    const uint32_t t = lds_u32(sched_task(r));
    mbar_arrive_cluster_relaxed(mapa(sched_empty_bar(r), 0), t);
    ++fetches;
    return (int)t;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs); the leader's is also the scheduler =====================
    // the whole warp walks the loop (uniform control flow and operands); lane 0 waits and issues
    uint32_t published = 0;
    auto publish = [&](uint32_t idx) -> int {       // leader lane 0: hand task index `idx` to every role of the pair
      const int t = idx < (uint32_t)ntasks ? (int)idx : -1;
      const int r = (int)(published % kSched);
      mbar_wait(sched_empty_bar(r), ((published / kSched) & 1) ^ 1);
#pragma unroll
      for (uint32_t cta = 0; cta < 2; ++cta) {
        const uint32_t bar = mapa(sched_full_bar(r), cta);
        mbar_arrive_expect_tx_cluster_relaxed(bar, 4u);
        st_async_u32(mapa(sched_task(r), cta), (uint32_t)t, bar);
      }
      ++published;
      return t;
    };
    // first task = the cluster's index (no atomic round trip at launch); the next ones come from the queue, popped at
    // the start of the running task and published once its first stages are in flight.  The static first task assumes
    // that every cluster becomes resident without waiting for another kernel: when a collective that waits on a peer
    // GPU may hold SMs (data parallel runs), `dynamic_first` makes the first task come from the queue as well, so that
    // the smallest unfinished task is always held by a RUNNING cluster whatever share of the SMs this launch gets.
    const int nclusters = (int)(gridDim.x >> 1);
    const uint32_t queue_base = dynamic_first ? 0u : (uint32_t)nclusters;
    int t = -1;
    if (lane == 0) {
      if (rank == 0) {
        const uint32_t first = dynamic_first ? atomicAdd(queue, 1u) : (uint32_t)(blockIdx.x >> 1);
        t = publish(first);
      } else {
        t = next_task();
      }
    }
    t = bcast(t);
    GTask tk = load_task(tasks, t);
    int ps = 0; uint32_t pphase = 0;       // ring position across tasks: stage, and the parity of its current use
    const uint32_t full_leader0 = mapa(full_bar(0), 0);
    while (t >= 0) {
      uint32_t raw = 0;
      if (lane == 0 && rank == 0) raw = atomicAdd(queue, 1u);      // consumed below, after the first loads are issued
      if (tk.flags & TF_ELT) {               // nothing to stream: hand the next task to every role at once
        int t_next = -1;
        if (elect_one()) t_next = (rank == 0) ? publish(queue_base + raw) : next_task();
        __syncwarp();
        t = bcast(t_next);
        tk = load_task(tasks, t);
        continue;
      }
      const GProblem* p = &params.p[tk.problem];
      const int BN = tk.bn, BNH = BN >> 1;
      const bool a_mn = (tk.flags & TF_A_MN) != 0, b_mn = (tk.flags & TF_B_MN) != 0;
      const int rpc = a_mn ? BM_CTA : rows_per_cta(tk.M, small_rows);   // (TN: M = features, always 128 per CTA)
      const int m0 = tk.m_blk * BM + (int)rank * rpc;         // this CTA's rows of A
      const int nb0 = tk.n_blk * BN + (int)rank * BNH;        // this CTA's slice of B
      const uint32_t stage_tx = (k64 ? 4u : 2u) * ((uint32_t)rpc * BK * 4 + (uint32_t)BNH * BK * 4);
      const bool half = (tk.flags & TF_HALF) != 0;
      const int W = half ? tk.wait2_val : tk.nkb;              // k-blocks per producing tile
      if (lane == 0) {
        if (half) {
          wait_counter(counters + tk.wait2_ctr, (uint32_t)tk.wait_val);    // first halves of the producing tiles
          fence_acq_rel_gpu();
          fence_proxy_async_all();
        } else if (tk.wait_cnt > 0 || tk.wait2_ctr >= 0) {
          for (int c = 0; c < tk.wait_cnt; ++c) wait_counter(counters + tk.wait_ctr + c, (uint32_t)tk.wait_val);
          if (tk.wait2_ctr >= 0) wait_counter(counters + tk.wait2_ctr, (uint32_t)tk.wait2_val);
          fence_acq_rel_gpu();               // pairs with the signal warps' fence + counter add
          fence_proxy_async_all();           // their generic-proxy stores -> our TMA (async proxy) loads
        }
        if (kTimeline && tl && rank == 0) { tl[kTL * t + 0] = gtimer(); tl[kTL * t + 5] = blockIdx.x >> 1; tl[kTL * t + 6] = t_entry; }
      }
      __syncwarp();
      int t_after = -1;
      const int announce = min(tk.nkb, nstages) - 1;
      int i = 0;
      for (int pass = 0; pass < (half ? 2 : 1); ++pass) {
        if (pass == 1) {                     // the rest of the producing tiles
          if (lane == 0) {
            for (int c = 0; c < tk.wait_cnt; ++c) wait_counter(counters + tk.wait_ctr + c, (uint32_t)tk.wait_val);
            fence_acq_rel_gpu();
            fence_proxy_async_all();
          }
          __syncwarp();
        }
        for (int j0 = 0; j0 < tk.nkb; j0 += W) {
          const int nch = min(W, tk.nkb - j0);
          const int hc = half ? half_chunks(nch) : nch;
          const int c_lo = pass == 0 ? 0 : hc, c_hi = pass == 0 ? hc : nch;
          for (int c = c_lo; c < c_hi; ++c, ++i) {
            const int s = ps;
            const uint32_t sa = base + (uint32_t)s * stage_bytes, sb = sa + b_off;
            const uint32_t full_leader = full_leader0 + 8u * s;
            const int k0 = (tk.kb0 + j0 + c) * (k64 ? 2 * BK : BK);
            if (elect_one()) {
              mbar_wait(empty_bar(s), pphase ^ 1);
              if (kTimeline && tl && t == 0 && i < 64) tl[kTL * ntasks + (rank ? 128 : 0) + i] = gtimer();
              if (rank == 0) mbar_arrive_expect_tx(full_bar(s), stage_tx);
              // MN-major operands: one 3-D box {32 mn, 32 k, chunks} lands as [chunk][k][32 mn] (see make_map_mn)
              // (64-deep k-blocks: K-major operands are 3-D {32 k, rows, k chunk}, box {32, rows, 2})
              if (a_mn) tma_load_3d_pair(sa, &p->map_a, full_leader, 0, k0, m0 >> 5);
              else if (k64) tma_load_3d_pair(sa, &p->map_a, full_leader, 0, m0, k0 >> 5);
              else tma_load_2d_pair(sa, &p->map_a, full_leader, k0, m0);
              if (b_mn) tma_load_3d_pair(sb, &p->map_b, full_leader, 0, k0, nb0 >> 5);
              else if (k64) tma_load_3d_pair(sb, &p->map_b, full_leader, 0, nb0, k0 >> 5);
              else tma_load_2d_pair(sb, &p->map_b, full_leader, k0, nb0);
              if (i == announce)               // the next task: known to every role while this one streams
                t_after = (rank == 0) ? publish(queue_base + raw) : next_task();
            }
            __syncwarp();
            if (++ps == nstages) { ps = 0; pphase ^= 1u; }
          }
        }
      }
      t = bcast(t_after);
      tk = load_task(tasks, t);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA): whole warp walks, lane 0 waits and issues =====================
    if (rank == 0) {
      uint32_t tcount = 0, mphase = 0;
      int ms = 0;                            // ring position across tasks (the producers' ps / pphase)
      int t = 0;
      if (lane == 0) t = next_task();
      t = bcast(t);
      GTask tk = load_task(tasks, t);
      for (; t >= 0; ++tcount) {
        int t_after = -1;
        if (tk.flags & TF_ELT) {             // no accumulator involved (tcount counts contraction tasks only)
          if (lane == 0) t_after = next_task();
          t = bcast(t_after);
          tk = load_task(tasks, t);
          --tcount;
          continue;
        }
        const int announce = min(tk.nkb, nstages) - 1;   // the producer publishes the next task at this k-block
        const bool a_mn = (tk.flags & TF_A_MN) != 0, b_mn = (tk.flags & TF_B_MN) != 0;
        const uint32_t idesc = make_idesc(BM, tk.bn, a_mn, b_mn);
        // 64-deep k-blocks: chunk 1 of a K-major operand starts behind the rows of chunk 0
        const uint32_t a_chunk = (uint32_t)(a_mn ? BM_CTA : rows_per_cta(tk.M, small_rows)) * 128u, b_chunk = (uint32_t)(tk.bn >> 1) * 128u;
        const uint32_t acc = tcount & 1;
        const uint32_t tmem_d = tmem_base + acc * kAccCols;
        if (lane == 0) {
          mbar_wait(tmem_empty_bar(acc), ((tcount >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
          tc_fence_after();
          if (kTimeline && tl) { tl[kTL * t + 1] = gtimer(); tl[kTL * t + 7] = clock64(); }
        }
        __syncwarp();
        for (int i = 0; i < tk.nkb; ++i) {
          const int s = ms;
          const uint32_t sa = base + (uint32_t)s * stage_bytes, sb = sa + b_off;
          if (elect_one()) {
            mbar_wait(full_bar(s), mphase);
            tc_fence_after();
            if (kTimeline && tl && t == 0 && i < 64) tl[kTL * ntasks + 64 + i] = gtimer();
            if (k64) {
#pragma unroll
              for (int k = 0; k < 2 * BK / UMMA_K; ++k) {
                const uint64_t da = a_mn ? desc_mn_major(sa + k * 1024, 2 * BK) : desc_k_major(sa + (k >> 2) * a_chunk + (k & 3) * 32);
                const uint64_t db = b_mn ? desc_mn_major(sb + k * 1024, 2 * BK) : desc_k_major(sb + (k >> 2) * b_chunk + (k & 3) * 32);
                umma_tf32_pair(tmem_d, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
              }
            } else {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k) {
                const uint64_t da = a_mn ? desc_mn_major(sa + k * 1024) : desc_k_major(sa + k * 32);
                const uint64_t db = b_mn ? desc_mn_major(sb + k * 1024) : desc_k_major(sb + k * 32);
                umma_tf32_pair(tmem_d, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
              }
            }
            umma_commit_pair(empty_bar(s));     // frees the smem slot in both CTAs once these MMAs have read it
            if (i == announce) t_after = next_task();
          }
          __syncwarp();
          if (++ms == nstages) { ms = 0; mphase ^= 1u; }
        }
        if (elect_one()) {
          umma_commit_pair(tmem_full_bar(acc)); // accumulator complete (both CTAs)
          if (kTimeline && tl) { tl[kTL * t + 2] = gtimer(); tl[kTL * t + 7] = clock64() - tl[kTL * t + 7]; }
        }
        t = bcast(t_after);
        tk = load_task(tasks, t);
      }
    }
  } else if (warp == kSignalWarp) {
    // ===================== signal warp (both CTAs): publishes finished tiles to the other clusters =====================
    // Making a tile's stores visible GPU-wide costs a MEMBAR.GPU (~0.9 us: every outstanding store of the SM must be
    // acknowledged by L2).  The epilogue warps used to pay it at the end of every tile (red.release.gpu), 5 % of their
    // time; now they only arrive on a CTA-local barrier and this warp releases: its acquire of the barrier makes their
    // stores cumulative with its fence.acq_rel.gpu, and the counter add carries the CTA's kEpiWarps arrivals at once.
    if (lane == 0) {
      uint32_t seq = 0;
      int t = next_task();
      while (t >= 0) {
        const GTask* tp = tasks + t;
        const int flags = __ldg(&tp->flags), sig = __ldg(&tp->signal_ctr);
        if (sig >= 0) {                    // tile tasks and elementwise tasks alike
          // (every signalling tile cycles both barriers of its ring slot, so that their phases stay in step)
          mbar_wait(half_bar(seq), (seq / kDoneRing) & 1);
          if (flags & TF_SIG_HALF) {
            fence_acq_rel_gpu();
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(counters + sig + half_off), "r"((uint32_t)kEpiWarps) : "memory");
          }
          mbar_wait(done_bar(seq), (seq / kDoneRing) & 1);
          fence_acq_rel_gpu();
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(counters + sig), "r"((uint32_t)kEpiWarps) : "memory");
          ++seq;
          asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(sig_seq_addr), "r"(seq) : "memory");
        }
        t = next_task();
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..9 of both CTAs) =====================
    const int e = warp - 2;                 // epilogue warp index
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int slot = e >> 2;                // this warp takes chunks slot, slot + kSlots, ...
    const uint32_t ebuf = epi_base + e * EPI_WARP_BYTES;         // kEpiBufs rotating 32 x 32 boxes
    const uint32_t tmem_empty_leader0 = mapa(tmem_empty_bar(0), 0), tmem_empty_leader1 = mapa(tmem_empty_bar(1), 0);
    uint32_t tcount = 0, cidx = 0, aux_phase = 0, red_pending = 0;   // cidx: chunks processed so far (box = cidx % 3)
    uint32_t sig_count = 0;                 // signalling tiles finished so far (index into the done-barrier ring)
    auto next_task_warp = [&]() -> int {
      int t = 0;
      if (lane == 0) t = next_task();
      return __shfl_sync(0xffffffffu, t, 0);
    };
    int t = next_task_warp();
    GTask tk = load_task(tasks, t);
    for (; t >= 0; ++tcount) {
      const int t_after = next_task_warp();
      const GTask tk_after = load_task(tasks, t_after);
      if (tk.flags & TF_ELT) {
        // ---- elementwise task: rows [256 m_blk + 128 rank, +128) of the batch, one thread per row (warps 0..3) ----
        if (kTimeline && tl && rank == 0 && warp == 2 && lane == 0) { tl[kTL * t + 3] = gtimer(); tl[kTL * t + 5] = blockIdx.x >> 1; tl[kTL * t + 6] = t_entry; }
        if (lane == 0) {
          for (int c = 0; c < tk.wait_cnt; ++c) wait_counter(counters + tk.wait_ctr + c, (uint32_t)tk.wait_val);
          if (tk.wait2_ctr >= 0)
            for (int c = 0; c < max(tk.kb0, 1); ++c) wait_counter(counters + tk.wait2_ctr + c, (uint32_t)tk.wait2_val);
          fence_acq_rel_gpu();               // pairs with the producers' red.release.gpu (inputs are read with ld.cg)
        }
        __syncwarp();
        if (kTimeline && tl && rank == 0 && warp == 2 && lane == 0) tl[kTL * t + 0] = gtimer();
        if (tk.flags & TF_ELT_FINALIZE) {
          // ---- cost finalize: one warp, lanes stride the block partials, fixed-order sums (finalize_kernel's arithmetic) ----
          if (e == 0 && rank == 0) {
            const FinalizeArgs& f = elem.fin;
            float sums[9];
#pragma unroll
            for (int qi = 0; qi < 9; ++qi) sums[qi] = 0.f;
            const int nq = 2 * f.n_mod + 1;
            for (int qi = 0; qi < nq; ++qi) {
              const float* src; int nblk, slot, stride = kCostSlots, off;
              if (qi == 2 * f.n_mod) { src = f.partials_latent; nblk = f.blocks_latent; slot = 8; off = 8; }
              else if (qi & 1) { src = f.partials_latent; nblk = f.blocks_latent; slot = qi; off = qi; }
              else { src = f.partials_recon[qi >> 1]; nblk = f.blocks_recon[qi >> 1]; slot = qi; stride = f.stride_recon[qi >> 1]; off = f.off_recon[qi >> 1]; }
              float acc = 0.f;
#pragma unroll 8
              for (int b = lane; b < nblk; b += 32) acc += __ldcg(src + (int64_t)b * stride + off);
              acc = warp_sum(acc);
#pragma unroll
              for (int k = 0; k < 9; ++k) if (k == slot) sums[k] = acc;
            }
            if (lane == 0) finalize_combine(f, sums, fin_advance);
          }
        } else if (e < 4) {
          const int64_t r = (int64_t)tk.m_blk * BM + (int64_t)rank * BM_CTA + e * 32 + lane;
          const bool live = r < (int64_t)tk.M;
          if (tk.flags & TF_ELT_LATENT_FWD) {
            float kl0 = 0.f, kl1 = 0.f, assoc = 0.f;
            if (live) {
              if (elem.lf.n_mod == 1) {
                float rk[1];
                latent_fwd_row<1, LoadCg>(elem.lf, r, rk, assoc, tk.act != 0);
                kl0 = rk[0];
              } else {
                float rk[2];
                latent_fwd_row<2, LoadCg>(elem.lf, r, rk, assoc, tk.act != 0);
                kl0 = rk[0]; kl1 = rk[1];
              }
            }
            kl0 = warp_sum(kl0); kl1 = warp_sum(kl1); assoc = warp_sum(assoc);
            if (lane == 0) {
              float* part = elem.lf.partials + (size_t)((tk.m_blk * 2 + (int)rank) * 4 + e) * kCostSlots;
              part[1] = kl0; part[3] = kl1; part[8] = assoc;
            }
          } else {
            const int nz = elem.lb.n_z;
            if (nz == 4) {
              for (int m = 0; m < elem.lb.n_mod; ++m) latent_bwd_row4<true>(elem.lb, m, r, live, elem.bh_grad[m], lane);
            } else
            for (int m = 0; m < elem.lb.n_mod; ++m) {
              for (int k = 0; k < nz; ++k) {
                float dm = 0.f, dl = 0.f;
                if (live) latent_bwd_elem<LoadCg>(elem.lb, m, r, k, dm, dl);
                if (elem.bh_grad[m] != nullptr) {
                  dm = warp_sum(dm); dl = warp_sum(dl);
                  if (lane == 0) { atomicAdd(elem.bh_grad[m] + k, dm); atomicAdd(elem.bh_grad[m] + nz + k, dl); }
                }
              }
            }
          }
        }
        __syncwarp();
        if (kTimeline && tl && rank == 0 && warp == 2 && lane == 0) tl[kTL * t + 8] = gtimer();
        // published like a tile: CTA-local arrival, the signal warp fences once and adds this CTA's kEpiWarps arrivals
        // (sixteen red.release.gpu of the warps themselves took ~3 us longer to reach the consumers)
        if (tk.signal_ctr >= 0) {
          if (elect_one()) {
            if (sig_count >= (uint32_t)kDoneRing) {
              uint32_t done;
              do { asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(done) : "r"(sig_seq_addr) : "memory"); } while (done + kDoneRing <= sig_count);
            }
            asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(half_bar(sig_count)) : "memory");
            asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(done_bar(sig_count)) : "memory");
          }
          ++sig_count;
        }
        if (kTimeline && tl && rank == 0 && warp == 2 && lane == 0) tl[kTL * t + 4] = gtimer();
        t = t_after;
        tk = tk_after;
        --tcount;
        continue;
      }
      const GProblem* p = &params.p[tk.problem];
      const int BN = tk.bn, N = tk.N, act = tk.act;
      const bool reduce = (tk.flags & TF_REDUCE) != 0, use_aux = (tk.flags & TF_AUX) != 0, round_out = (tk.flags & TF_ROUND) != 0;
      const float* __restrict__ bias = p->bias;
      const bool loss_task = (tk.flags & TF_LOSS) != 0;
      const int sig_hc = (tk.flags & TF_SIG_HALF) ? half_chunks(min(tk.bn / 32, (tk.N - tk.n_blk * tk.bn + 31) / 32)) : 0;
      float loss_acc = 0.0f;                 // this lane's row: reconstruction loss over the warp's chunks
      float* __restrict__ colsum = (tk.flags & TF_COLSUM) ? p->colsum : nullptr;
      const int rpc = (tk.flags & TF_A_MN) ? BM_CTA : rows_per_cta(tk.M, small_rows);
      const int row0 = tk.m_blk * BM + (int)rank * rpc + q * 32;       // first output row of this warp
      const int m_lim = min(tk.M, tk.m_blk * BM + (int)rank * rpc + rpc);   // rows of this CTA end here (warp-uniform)
      const int n0 = tk.n_blk * BN;
      const int nchunks = (row0 < m_lim) ? min(BN / 32, (N - n0 + 31) / 32) : 0;   // warp-uniform
      const int nmine = nchunks > slot ? (nchunks - slot + kSlots - 1) / kSlots : 0;   // chunks slot, slot + kSlots, ... of this warp
      const uint32_t acc = tcount & 1;
      if (sig_hc > 0 && (nmine == 0 || slot >= sig_hc)) {      // no first-half chunk of this warp: nothing to wait for
        if (elect_one()) {
          if (sig_count >= (uint32_t)kDoneRing) {
            uint32_t done;
            do { asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(done) : "r"(sig_seq_addr) : "memory"); } while (done + kDoneRing <= sig_count);
          }
          asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(half_bar(sig_count)) : "memory");
        }
      }
      // everything the chunk loop needs from global memory is requested NOW, while the main loop of this task runs
      // (a fresh L2 round trip costs ~1.5 us while the operand streams of 148 SMs are in flight): the bias values of all
      // chunks of this warp (lane <-> column), the relu mask words of this lane's row, and up to three aux boxes
      const uint32_t* __restrict__ mask_in = (tk.flags & TF_MASK_IN) ? p->mask_in : nullptr;
      uint32_t* __restrict__ mask_out = (tk.flags & TF_MASK_OUT) ? p->mask_out : nullptr;
      float bv0 = 0.0f, bv1 = 0.0f, bv2 = 0.0f, bv3 = 0.0f;
      if (bias != nullptr) {
        const int col = n0 + slot * 32 + lane;
        constexpr int cs = 32 * kSlots;                    // column stride between this warp's chunks
        if (0 < nmine && col < N) bv0 = __ldg(bias + col);
        if (1 < nmine && col + cs < N) bv1 = __ldg(bias + col + cs);
        if (2 < nmine && col + 2 * cs < N) bv2 = __ldg(bias + col + 2 * cs);
        if (3 < nmine && col + 3 * cs < N) bv3 = __ldg(bias + col + 3 * cs);
      }
      uint32_t mw0 = 0u, mw1 = 0u, mw2 = 0u, mw3 = 0u;
      if (mask_in != nullptr && row0 + lane < m_lim) {
        const uint32_t* mrow = mask_in + (size_t)(row0 + lane) * (size_t)p->ldmask + (n0 >> 5) + slot;
        if (0 < nmine) mw0 = __ldg(mrow);
        if (1 < nmine) mw1 = __ldg(mrow + kSlots);
        if (2 < nmine) mw2 = __ldg(mrow + 2 * kSlots);
        if (3 < nmine) mw3 = __ldg(mrow + 3 * kSlots);
      }
      if (use_aux && nmine > 0 && elect_one()) {
        bulk_wait_read<0>();               // a reduce-add of the previous task may still be reading these boxes
#pragma unroll
        for (int i = 0; i < kEpiBufs; ++i) {
          if (i < nmine) {
            const uint32_t b = (cidx + i) % kEpiBufs;
            mbar_arrive_expect_tx(aux_bar(e, b), CHUNK_BYTES);
            tma_load_2d(ebuf + b * CHUNK_BYTES, &p->map_aux, aux_bar(e, b), n0 + (slot + kSlots * i) * 32, row0);
          }
        }
      }
      mbar_wait(tmem_full_bar(acc), (tcount >> 1) & 1);
      tc_fence_after();
      if (kTimeline && tl && rank == 0 && warp == 2 && lane == 0) { tl[kTL * t + 3] = gtimer(); tl[kTL * t + 15] = (unsigned long long)clock64(); }
      // a bulk reduce-add of an earlier task may still be reading one of the boxes: drain those reads once, here,
      // instead of polling in every chunk (reduce tasks themselves keep the per-chunk wait below)
      if (!reduce && !tma_store && red_pending) {
        if (!use_aux && elect_one()) bulk_wait_read<0>();      // (the aux branch above has already waited)
        __syncwarp();
        red_pending = 0;
      }
      const bool interior_rows = row0 + 32 <= m_lim;
      const uint32_t sts_base = ebuf + (uint32_t)lane * 128u, sts_x = (uint32_t)(lane & 7);
      const int rr = lane >> 3, jj = lane & 7;
      // read-back offsets of the transposed copy: row rr + 4k, 16-byte column jj; (rr + 4k) & 7 = rr + 4 (k & 1)
      const uint32_t rb_even = (uint32_t)rr * 128u + ((uint32_t)(jj ^ rr) << 4);
      const uint32_t rb_odd = (uint32_t)rr * 128u + ((uint32_t)(jj ^ (rr + 4)) << 4);
      // accumulator bits of this lane's row, transformed in place.  The TMEM read of chunk i + 1 is issued as soon as
      // chunk i sits in shared memory, so that it overlaps the copy-out of chunk i (8 warps reading TMEM at once get
      // ~32 B/clk: ~1000 cycles per 4 KB chunk)
      uint32_t v[32];
      const uint32_t tmem_row = tmem_base + acc * kAccCols + ((uint32_t)(q * 32) << 16);
      if (stagger_ns != 0u && (slot & 1) && nmine > 0) __nanosleep(stagger_ns);
      if (nmine > 0) tmem_ld32_issue(tmem_row + (uint32_t)(slot * 32), v);
#pragma unroll 1
      for (int i = 0; i < nmine; ++i, ++cidx) {
        const int c = slot + kSlots * i;
        const uint32_t b = cidx % kEpiBufs;
        const uint32_t ob = ebuf + b * CHUNK_BYTES;
        tmem_ld_wait(v);
        if (kTimeline && tl && rank == 0 && warp == 2 && lane == 0 && i == 0) tl[kTL * t + 9] = (unsigned long long)clock64();
        if (dbg_skip_math) {
        } else if (mask_in != nullptr) {
          // relu': one bit per element, already in registers; the tf32 rounding (round-to-nearest, ties away = add half
          // an ulp to the magnitude, truncate) is folded into the same AND
          const uint32_t mw = i == 0 ? mw0 : i == 1 ? mw1 : i == 2 ? mw2 : mw3;
          if (round_out) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const uint32_t keep = (uint32_t)((int32_t)(mw << (31 - j)) >> 31);
              v[j] = (v[j] + 0x1000u) & (keep & 0xffffe000u);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] &= (uint32_t)((int32_t)(mw << (31 - j)) >> 31);
          }
        } else if (loss_task) {
          // output layer of a decoder: pre-activation -> reconstruction loss (:321-328) and d cost / d pre-activation;
          // the aux box holds the target tile x (rows / columns past M / N zero-filled by TMA and masked here)
          mbar_wait(aux_bar(e, b), (aux_phase >> b) & 1u);
          aux_phase ^= 1u << b;
          if (bias != nullptr) {
            const float b_cur = i == 0 ? bv0 : i == 1 ? bv1 : i == 2 ? bv2 : bv3;
            const uint32_t strip = strip_base + (uint32_t)e * kBiasStrip;
            __syncwarp();
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(strip + (uint32_t)lane * 4u), "f"(b_cur) : "memory");
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 bq = lds128(strip + (uint32_t)j * 16u);
              v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + bq.x);
              v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + bq.y);
              v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + bq.z);
              v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + bq.w);
            }
          }
          const int ncols = (row0 + lane < m_lim) ? min(32, N - (n0 + c * 32)) : 0;    // live columns of this lane's row
          const uint32_t xrow = ob + (uint32_t)lane * 128u, xsw = (uint32_t)(lane & 7);
          // (the four variants are separate straight-line loops: a per-element select of the loss form serialised the 32
          // independent MUFU chains of a chunk behind branches -- 18 us per tile instead of 7)
          if (p->loss_binary) {
            if (ncols == 32) loss_acc += loss_chunk<true, true>(v, xrow, xsw, p->loss_scale, round_out, 32);
            else loss_acc += loss_chunk<true, false>(v, xrow, xsw, p->loss_scale, round_out, ncols);
          } else {
            if (ncols == 32) loss_acc += loss_chunk<false, true>(v, xrow, xsw, p->loss_scale, round_out, 32);
            else loss_acc += loss_chunk<false, false>(v, xrow, xsw, p->loss_scale, round_out, ncols);
          }
        } else if (use_aux) {
          mbar_wait(aux_bar(e, b), (aux_phase >> b) & 1u);
          aux_phase ^= 1u << b;
          float h[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 hv = lds128(ob + swz(lane, j));
            h[4 * j] = hv.x; h[4 * j + 1] = hv.y; h[4 * j + 2] = hv.z; h[4 * j + 3] = hv.w;
          }
          if (act == ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = h[j] > 0.0f ? v[j] : 0u;
          } else if (act == ACT_SOFTPLUS) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * (1.0f - __expf(-h[j])));
          } else if (act == ACT_SIGMOID) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * (h[j] * (1.0f - h[j])));
          }
          if (round_out) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (v[j] + 0x1000u) & 0xffffe000u;
          }
        } else if (!reduce) {
          if (bias != nullptr) {
            const float b_cur = i == 0 ? bv0 : i == 1 ? bv1 : i == 2 ? bv2 : bv3;
            if (bias_smem) {
              // 32 bias values of the chunk through a 128-byte strip: 1 store + 8 broadcast 128-bit loads instead of 32 shuffles
              const uint32_t strip = strip_base + (uint32_t)e * kBiasStrip;
              __syncwarp();
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(strip + (uint32_t)lane * 4u), "f"(b_cur) : "memory");
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 bq = lds128(strip + (uint32_t)j * 16u);
                v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + bq.x);
                v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + bq.y);
                v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + bq.z);
                v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + bq.w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __shfl_sync(0xffffffffu, b_cur, j));
            }
          }
          if (act == ACT_RELU) {
            // on the bits: negative floats are negative integers, so max(., 0) is the relu; with rounding, the half ulp
            // is added first (a negative value stays negative, +0 .. half an ulp round to 0)
            if (round_out) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = (uint32_t)max((int32_t)(v[j] + 0x1000u), 0) & 0xffffe000u;
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = (uint32_t)max((int32_t)v[j], 0);
            }
          } else {
            if (act == ACT_SOFTPLUS) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float xv = __uint_as_float(v[j]);
                v[j] = __float_as_uint(fmaxf(xv, 0.0f) + __logf(1.0f + __expf(-fabsf(xv))));
              }
            } else if (act == ACT_SIGMOID) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__fdividef(1.0f, 1.0f + __expf(-__uint_as_float(v[j]))));
            }
            if (round_out) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = (v[j] + 0x1000u) & 0xffffe000u;
            }
          }
        }
        const bool interior_cols = n0 + c * 32 + 32 <= N;
        if (!interior_cols) {              // the pad columns N..roundup4(N) of the row pitch receive zeros
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (n0 + c * 32 + j < N) ? v[j] : 0u;
        }
        if (kTimeline && tl && rank == 0 && warp == 2 && lane == 0 && i == 0) tl[kTL * t + 11] = (unsigned long long)clock64();
        if (mask_out != nullptr && !dbg_skip_math) {
          // bit j = (output j > 0); relu outputs are >= +0, so "bits != 0": sign of the negated bits, shifted in from
          // the right, two independent chains of 16
          uint32_t wl = 0u, wh = 0u;
#pragma unroll
          for (int j = 15; j >= 0; --j) {
            wl = __funnelshift_l(0u - v[j], wl, 1);
            wh = __funnelshift_l(0u - v[j + 16], wh, 1);
          }
          if (interior_rows || row0 + lane < m_lim)
            mask_out[(size_t)(row0 + lane) * (size_t)p->ldmask + (n0 >> 5) + c] = (wh << 16) | wl;
        }
        if (kTimeline && tl && rank == 0 && warp == 2 && lane == 0 && i == 0) tl[kTL * t + 12] = (unsigned long long)clock64();
        if (rpc != BM_CTA && !interior_rows && row0 + lane >= m_lim) {
          // split rows: A rows at or past rows_per_cta were not loaded (the stage holds an earlier task's data), so this
          // lane's accumulator row is garbage -- the column sums and the TMA reduce-add read the staged tile
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        if (reduce || tma_store) {
          // box b was handed to a bulk store / reduce-add three chunks ago: wait until that one has read it
          if (elect_one()) bulk_wait_read<kEpiBufs - 1>();
          __syncwarp();
        }
        if (kTimeline && tl && rank == 0 && warp == 2 && lane == 0 && i == 0) tl[kTL * t + 13] = (unsigned long long)clock64();
        // (in the aux case the lane overwrites exactly the 128 bytes it has just read: in place, no hazard)
        {
          const uint32_t sb = sts_base + b * CHUNK_BYTES;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sb + (((uint32_t)j ^ sts_x) << 4)), "r"(v[4 * j]),
                         "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
        }
        if (kTimeline && tl && rank == 0 && warp == 2 && lane == 0 && i == 0) tl[kTL * t + 14] = (unsigned long long)clock64();
        if (i + 1 < nmine) tmem_ld32_issue(tmem_row + (uint32_t)((c + kSlots) * 32), v);
        if (reduce || tma_store) {
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            if (reduce) tma_reduce_add_2d(&p->map_c, ob, n0 + c * 32, row0);
            else tma_store_2d(&p->map_c, ob, n0 + c * 32, row0);
            bulk_commit();
          }
          red_pending = 1;
        } else if (!dbg_skip_store) {
          // transposed read-back: 8 lanes cover one 128-byte row segment, 4 rows per instruction -> coalesced 128-bit
          // stores; rows >= M and columns >= roundup4(N) are clipped (the TMA loads zero-filled them)
          __syncwarp();
          const int col = n0 + c * 32 + jj * 4;
          float* dst = p->c_ptr + (size_t)(row0 + rr) * (size_t)p->ldc + col;
          const size_t step = 4 * (size_t)p->ldc;
          if (interior_rows && interior_cols) {
            float4 o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = lds128(ob + ((k & 1) ? rb_odd : rb_even) + (uint32_t)k * 512u);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + k * step), "f"(o[k].x), "f"(o[k].y), "f"(o[k].z), "f"(o[k].w) : "memory");
          } else if (col < ((N + 3) & ~3)) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (row0 + rr + 4 * k < m_lim) {
                const float4 o = lds128(ob + ((k & 1) ? rb_odd : rb_even) + (uint32_t)k * 512u);
                asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + k * step), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
              }
            }
          }
        }
        if (colsum != nullptr) {
          // bias gradient of the layer below = column sums of this dgrad output: lane <-> column, 32 conflict-free
          // reads of the staged tile (rows past M hold exact zeros), one fp32 RED per column and warp
          float cs0 = 0.0f, cs1 = 0.0f, cs2 = 0.0f, cs3 = 0.0f;
          const uint32_t cb = ob + (uint32_t)(lane & 3) * 4u;
          const uint32_t cj = (uint32_t)(lane >> 2);
#pragma unroll
          for (int r = 0; r < 32; r += 4) {
            float a0, a1, a2, a3;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a0) : "r"(cb + (uint32_t)r * 128u + ((cj ^ (uint32_t)(r & 7)) << 4)));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a1) : "r"(cb + (uint32_t)(r + 1) * 128u + ((cj ^ (uint32_t)((r + 1) & 7)) << 4)));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a2) : "r"(cb + (uint32_t)(r + 2) * 128u + ((cj ^ (uint32_t)((r + 2) & 7)) << 4)));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a3) : "r"(cb + (uint32_t)(r + 3) * 128u + ((cj ^ (uint32_t)((r + 3) & 7)) << 4)));
            cs0 += a0; cs1 += a1; cs2 += a2; cs3 += a3;
          }
          const int col = n0 + c * 32 + lane;
          if (col < N) atomicAdd(colsum + col, (cs0 + cs1) + (cs2 + cs3));
        }
        if (kTimeline && tl && rank == 0 && warp == 2 && lane == 0 && i == 0) tl[kTL * t + 10] = (unsigned long long)clock64();
        if (c < sig_hc && c + kSlots >= sig_hc) {
          // this warp's last chunk of the tile's first half is on its way to L2: let the signal warp publish the half
          __syncwarp();
          if (elect_one()) {
            if (sig_count >= (uint32_t)kDoneRing) {
              uint32_t done;
              do { asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(done) : "r"(sig_seq_addr) : "memory"); } while (done + kDoneRing <= sig_count);
            }
            asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(half_bar(sig_count)) : "memory");
          }
        }
        if (use_aux && i + kEpiBufs < nmine) {
          __syncwarp();                    // every lane has finished with box b: refill it with the aux tile 3 chunks ahead
          if (elect_one()) {
            mbar_arrive_expect_tx(aux_bar(e, b), CHUNK_BYTES);
            tma_load_2d(ob, &p->map_aux, aux_bar(e, b), n0 + (c + kSlots * kEpiBufs) * 32, row0);
          }
        }
      }
      // every tcgen05.ld of this accumulator has completed (wait::ld): hand it back to the MMA issuer; the global
      // stores of all lanes are ordered before the elected lane's release by the warp barrier
      if (kTimeline && tl && rank == 0 && warp == 2 && lane == 0) tl[kTL * t + 8] = gtimer();
      if (loss_task) {
        loss_acc = warp_sum(loss_acc);
        const int tiles_n = (N + BN - 1) / BN;
        if (lane == 0)
          p->loss_partials[(size_t)((tk.m_blk * tiles_n + tk.n_blk) * 2 + (int)rank) * kEpiWarps + e] = loss_acc;
      }
      tc_fence_before();
      __syncwarp();
      if (elect_one()) {
        mbar_arrive_cluster_relaxed(acc ? tmem_empty_leader1 : tmem_empty_leader0, 0u);
        if (tk.signal_ctr >= 0) {
          if (reduce || tma_store) { bulk_wait_complete(); fence_proxy_async_all(); }   // bulk stores performed, then publish
          // the signal warp must have consumed this ring slot's previous use (it lags by microseconds at most)
          if (sig_count >= (uint32_t)kDoneRing) {
            uint32_t done;
            do { asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(done) : "r"(sig_seq_addr) : "memory"); } while (done + kDoneRing <= sig_count);
          }
          if (sig_hc == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(half_bar(sig_count)) : "memory");
          asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(done_bar(sig_count)) : "memory");
        }
      }
      if (tk.signal_ctr >= 0) ++sig_count;
      if (kTimeline && tl && rank == 0 && warp == 2 && lane == 0) tl[kTL * t + 4] = gtimer();
      t = t_after;
      tk = tk_after;
    }
    if (elect_one()) bulk_wait_read<0>();   // shared memory must outlive the reads of the bulk stores
  }
  tc_fence_before();
  cluster_sync_all();     // the peer's smem / barriers stay valid until every MMA and every TMA of the pair is done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
  // self-cleaning: the last cluster to leave rewinds the queue and the row-block counters this launch used, so the
  // next launch (next step / next graph replay) needs no memset
  __shared__ uint32_t s_last;
  if (threadIdx.x == 0) {
    uint32_t last = 0;
    if (rank == 0) {
      __threadfence();
      last = (atomicAdd(queue + 1, 1u) == (gridDim.x >> 1) - 1) ? 1u : 0u;
    }
    s_last = last;
  }
  __syncthreads();
  if (s_last) {
    for (int i = threadIdx.x; i < reset_count; i += kThreads) {
      counters[reset_first + i] = 0u;
      if (half_off > 0) counters[reset_first + half_off + i] = 0u;
    }
    if (threadIdx.x == 0) { queue[0] = 0u; queue[1] = 0u; }
    __threadfence();
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 tensor map: dim0 (contiguous) x dim1 with row pitch `ld` floats, box {32, box_rows}
bool make_map(CUtensorMap* map, const float* ptr, int64_t dim0, int64_t dim1, int64_t ld, int box_rows, bool atom32,
              char* err, int errlen) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { snprintf(err, errlen, "cuTensorMapEncodeTiled entry point not available"); return false; }
  cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled failed (%d) ptr=%p dims=%lld x %lld ld=%lld box_rows=%d", (int)r,
             (const void*)ptr, (long long)dim0, (long long)dim1, (long long)ld, box_rows);
    return false;
  }
  return true;
}

// K-major operand [rows, K contiguous] for 64-deep k-blocks: a 3-D tensor {32 k, rows, ceil(K / 32) chunks} with strides
// {4 B, pitch, 128 B}; ONE box {32, box_rows, 2} lands as [chunk][row][32 k] = two ordinary K-major tiles.  The last
// chunk's k >= K - 32 c reads the pad columns of the row (pitches of the dense modalities are multiples of 32 floats; z and
// the heads gradient, whose pitch is their width, read into the following rows / the buffer's guard region): finite
// values, multiplied by the other operand's zeros -- its own k rows >= K are out of bounds (MN-major: zero-filled) or pad
// columns that stay exactly 0 (weights under Adam).  Chunks past the end are zero-filled.
bool make_map_k64(CUtensorMap* map, const float* ptr, int64_t dim_k, int64_t dim_rows, int64_t ld, int box_rows, char* err,
                  int errlen) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { snprintf(err, errlen, "cuTensorMapEncodeTiled entry point not available"); return false; }
  // (rows narrower than a chunk -- z: pitch n_z, the heads gradient: 2 n_z -- expose only their pitch; the box's remaining k
  // are out of bounds and zero-filled, and no two rows of the map overlap)
  cuuint64_t dims[3] = {(cuuint64_t)std::min<int64_t>(32, ld), (cuuint64_t)dim_rows, (cuuint64_t)((dim_k + 31) / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128u};
  cuuint32_t box[3] = {32u, (cuuint32_t)box_rows, 2u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled (3-D, K-major) failed (%d) ptr=%p k=%lld rows=%lld ld=%lld box_rows=%d", (int)r,
             (const void*)ptr, (long long)dim_k, (long long)dim_rows, (long long)ld, box_rows);
    return false;
  }
  return true;
}

// MN-major operand [dim_k rows, dim_mn contiguous] as a 3-D tensor {32, dim_k, ceil(dim_mn / 32)}: element (x, k, c) =
// ptr[k * ld + 32 c + x].  The last chunk's x >= dim_mn - 32 c reads the first floats of the next row (or up to 124 B
// past the last row: every buffer of the library carries that slack); those lanes only feed output rows / columns
// past M / N, which the epilogue masks and the TMA store clips.  Rows k >= dim_k and chunks past the end are zero-filled.
bool make_map_mn(CUtensorMap* map, const float* ptr, int64_t dim_mn, int64_t dim_k, int64_t ld, int chunks, char* err,
                 int errlen, int box_k = 32) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { snprintf(err, errlen, "cuTensorMapEncodeTiled entry point not available"); return false; }
  cuuint64_t dims[3] = {32u, (cuuint64_t)dim_k, (cuuint64_t)((dim_mn + 31) / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128u};
  cuuint32_t box[3] = {32u, (cuuint32_t)box_k, (cuuint32_t)chunks};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(err, errlen, "cuTensorMapEncodeTiled (3-D) failed (%d) ptr=%p mn=%lld k=%lld ld=%lld chunks=%d", (int)r,
             (const void*)ptr, (long long)dim_mn, (long long)dim_k, (long long)ld, chunks);
    return false;
  }
  return true;
}

}  // namespace

// ---- group plan: the problems / tasks of a handle, cut into launch sites ---------------------------------------------
constexpr int kSiteProblemsSmall = 2, kSiteProblemsLarge = 24, kSiteProblemsStep = 36;

struct GroupSite { int first_problem = 0, n_problems = 0, first_task = 0, n_tasks = 0; };

struct GroupPlan {
  std::vector<GProblem> problems;
  std::vector<GTask> tasks;          // .problem is relative to the site's first problem once the site is closed
  std::vector<GroupSite> sites;
  GElem elem;                        // arguments of the elementwise tasks (one latent stage per handle)
  bool open = false;
  GTask* d_tasks = nullptr;
  uint32_t* d_counters = nullptr;    // row-block completion counters of the step (not owned; self-cleaning)
  int n_counters = 0;
  int half_off = 0;                  // offset of the half-tile counters (0: none)
  bool allow_k64 = true;             // 64-deep k-blocks permitted for this plan's problems (group_set_deep_k)
  bool uploaded = false;
};

GroupPlan* group_create() { return new GroupPlan(); }

void group_destroy(GroupPlan* g) {
  if (!g) return;
  if (g->d_tasks) cudaFree(g->d_tasks);
  delete g;
}

// opens a launch site: the problems / tasks added until group_end() run as ONE kernel launch
int group_begin(GroupPlan* g) {
  GroupSite st;
  st.first_problem = (int)g->problems.size();
  st.first_task = (int)g->tasks.size();
  g->sites.push_back(st);
  g->open = true;
  return (int)g->sites.size() - 1;
}

bool group_end(GroupPlan* g, char* err, int errlen) {
  GroupSite& st = g->sites.back();
  st.n_problems = (int)g->problems.size() - st.first_problem;
  st.n_tasks = (int)g->tasks.size() - st.first_task;
  g->open = false;
  if (st.n_problems > kSiteProblemsStep) {
    snprintf(err, errlen, "a launch site holds %d contractions, at most %d fit the kernel parameter", st.n_problems, kSiteProblemsStep);
    return false;
  }
  for (int i = 0; i < st.n_tasks; ++i) g->tasks[st.first_task + i].problem -= st.first_problem;
  return true;
}

int group_site_tasks(const GroupPlan* g, int site) { return g->sites[site].n_tasks; }

bool tc_supported(int kind, const GemmArgs& a) {
  if (!a.A || !a.B || !a.C) return false;
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15) ||
      (reinterpret_cast<uintptr_t>(a.C) & 15))
    return false;
  if ((a.lda & 3) || (a.ldb & 3) || (a.ldc & 3)) return false;
  if (a.aux && ((reinterpret_cast<uintptr_t>(a.aux) & 15) || (a.ldaux & 3))) return false;
  if (a.bias && (reinterpret_cast<uintptr_t>(a.bias) & 15)) return false;
  // every contraction whose batch extent fills a tile row block runs here, including the n_z-wide ones (heads,
  // decoder input layer): those are HBM-bound and the TMA pipeline streams the one large operand exactly once;
  // out-of-range rows / columns of the narrow operand are zero-filled by TMA at no HBM cost
  const int batch = (kind == 2) ? a.K : a.M;
  return batch >= 32;
}

// tile width of a contraction with `batch_rows` batch rows (M of the row-wise forms, K of the weight gradient).  With
// many row blocks the widest tile wins (fewest operand bytes per FLOP: the main loops are bound by the L2 -> SM fabric).
// With few row blocks (B = 100: ONE) a 256-wide tile leaves 70 of the 74 CTA pairs idle and its epilogue (4 chunks per
// warp, ~5.5 us) sits on the layer-to-layer critical path: narrower tiles spread the columns over more pairs and cut the
// epilogue to 1-2 chunks per warp.
int group_tile_width(int N, int batch_rows) {
  const int row_blocks = (batch_rows + BM - 1) / BM;
  int max_bn = row_blocks >= 16 ? 256 : (row_blocks >= 4 ? 128 : 64);
  if (const char* e = getenv("VAEASSOC_MAX_BN")) { const int v = atoi(e); if (v == 64 || v == 128 || v == 192 || v == 256) max_bn = v; }
  const int tiles_n = (N + max_bn - 1) / max_bn;
  return std::min(max_bn, (((N + tiles_n - 1) / tiles_n) + 63) / 64 * 64);
}

bool deep_k_enabled() {
  static const bool on = getenv("VAEASSOC_NO_DEEP_K") == nullptr && getenv("VAEASSOC_MAX_BN") == nullptr &&
                         getenv("VAEASSOC_WIDE_TILES") == nullptr && getenv("VAEASSOC_RING_FIXED") == nullptr &&
                         getenv("VAEASSOC_PITCH_ALIGN") == nullptr;
  return on;
}
bool small_rows_enabled() {
  static const bool on = getenv("VAEASSOC_NO_SMALL_ROWS") == nullptr;
  return on;
}

// adds the contraction to the plan; returns its problem index or -1 (err filled)
int group_add_problem(GroupPlan* g, int kind, const GemmArgs& a, char* err, int errlen) {
  GProblem p;
  memset(&p, 0, sizeof p);
  const int BN = getenv("VAEASSOC_WIDE_TILES") ? group_tile_width(a.N, 1 << 20) : group_tile_width(a.N, kind == 2 ? a.K : a.M);
  bool ok = true;
  // 64-deep k-blocks where every tile of the plan is narrow, i.e. at batches of one to three row blocks (group_tile_width)
  const int batch_rows = kind == 2 ? a.K : a.M;
  // ... and where the K-major operands' chunk reads past K stay inside the row's zero pad (pitch a multiple of 32
  // floats: every activation / weight of the dense modalities) or are cut by the map itself (pitch below 32: z, d heads)
  auto k_major_ok = [](int64_t ld) { return ld % 32 == 0 || ld <= 32; };
  const bool k64 = deep_k_enabled() && g->allow_k64 && BN <= 64 && batch_rows <= 3 * BM &&
                   (kind == 2 || (k_major_ok(a.lda) && (kind == 0 || k_major_ok(a.ldb))));
  const int rpc = rows_per_cta(a.M, small_rows_enabled());
  const int bk = k64 ? 64 : 32;
  switch (kind) {
    case 0:    // NN: A [M,K] K-major ; B [K,N] MN-major
      ok = (k64 ? make_map_k64(&p.map_a, a.A, a.K, a.M, a.lda, rpc, err, errlen)
                : make_map(&p.map_a, a.A, a.K, a.M, a.lda, rpc, false, err, errlen)) &&
           make_map_mn(&p.map_b, a.B, a.N, a.K, a.ldb, BN / 64, err, errlen, bk);
      p.a_mn = 0; p.b_mn = 1;
      break;
    case 1:    // NT: A [M,K] K-major ; B [N,K] K-major
      ok = (k64 ? make_map_k64(&p.map_a, a.A, a.K, a.M, a.lda, rpc, err, errlen)
                : make_map(&p.map_a, a.A, a.K, a.M, a.lda, rpc, false, err, errlen)) &&
           (k64 ? make_map_k64(&p.map_b, a.B, a.K, a.N, a.ldb, BN / 2, err, errlen)
                : make_map(&p.map_b, a.B, a.K, a.N, a.ldb, BN / 2, false, err, errlen));
      p.a_mn = 0; p.b_mn = 0;
      break;
    default:   // TN: A [K,M] MN-major ; B [K,N] MN-major
      ok = make_map_mn(&p.map_a, a.A, a.M, a.K, a.lda, BM_CTA / 32, err, errlen, bk) &&
           make_map_mn(&p.map_b, a.B, a.N, a.K, a.ldb, BN / 64, err, errlen, bk);
      p.a_mn = 1; p.b_mn = 1;
      break;
  }
  p.k64 = k64 ? 1 : 0;
  ok = ok && make_map(&p.map_c, a.C, a.N, a.M, a.ldc, 32, false, err, errlen);
  const bool mask_in = kind != 2 && a.mask_in != nullptr && a.act == ACT_RELU && a.aux != nullptr;
  const bool loss = kind == 0 && a.loss_x != nullptr && a.loss_partials != nullptr;
  const bool has_aux = (kind != 2 && a.aux != nullptr && !mask_in) || loss;
  if (ok && loss) ok = make_map(&p.map_aux, a.loss_x, a.N, a.M, a.ld_loss_x, 32, false, err, errlen);
  else if (ok && has_aux) ok = make_map(&p.map_aux, a.aux, a.N, a.M, a.ldaux, 32, false, err, errlen);
  else if (ok) p.map_aux = p.map_c;
  if (!ok) return -1;
  p.M = a.M; p.N = a.N; p.K = a.K; p.BN = BN;
  p.reduce = (kind == 2 || a.force_reduce) ? 1 : 0;
  p.act = a.act; p.round_out = a.round_out; p.has_aux = has_aux ? 1 : 0;
  p.bias = (kind == 0 && !a.force_reduce) ? a.bias : nullptr;
  p.colsum = (kind != 2 && !a.force_reduce) ? a.bias_grad : nullptr;   // NN / NT: bias_grad = where the column sums of C go
  p.c_ptr = a.C; p.ldc = a.ldc;
  p.mask_in = mask_in ? a.mask_in : nullptr;
  p.mask_out = (kind == 0 && a.act == ACT_RELU && !a.aux) ? a.mask_out : nullptr;
  p.ldmask = a.ldmask;
  p.loss_partials = loss ? a.loss_partials : nullptr;
  p.loss_scale = a.loss_scale; p.loss_binary = a.loss_binary;
  g->problems.push_back(p);
  g->uploaded = false;
  return (int)g->problems.size() - 1;
}

int group_problem_tiles_m(const GroupPlan* g, int prob) { return (g->problems[prob].M + BM - 1) / BM; }
int group_problem_tiles_n(const GroupPlan* g, int prob) { const GProblem& p = g->problems[prob]; return (p.N + p.BN - 1) / p.BN; }
int group_problem_kblocks(const GroupPlan* g, int prob) {
  const int bk = g->problems[prob].k64 ? 2 * BK : BK;
  return (g->problems[prob].K + bk - 1) / bk;
}
int group_problem_kb_per_rowblock(const GroupPlan* g, int prob) { return g->problems[prob].k64 ? BM / (2 * BK) : BM / BK; }

int group_add_task(GroupPlan* g, int prob, int m_blk, int n_blk, int kb0, int nkb, int wait_ctr, int wait_cnt,
                   int wait_val, int wait2_ctr, int wait2_val, int signal_ctr, int extra_flags) {
  const GProblem& p = g->problems[prob];
  GTask t;
  memset(&t, 0, sizeof t);
  t.problem = prob; t.m_blk = m_blk; t.n_blk = n_blk; t.kb0 = kb0; t.nkb = nkb;
  t.wait_ctr = wait_ctr; t.wait_cnt = wait_cnt; t.wait_val = wait_val;
  t.wait2_ctr = wait2_ctr; t.wait2_val = wait2_val; t.signal_ctr = signal_ctr;
  t.bn = p.BN;
  t.flags = (p.a_mn ? TF_A_MN : 0) | (p.b_mn ? TF_B_MN : 0) | (p.reduce ? TF_REDUCE : 0) | (p.has_aux ? TF_AUX : 0) |
            (p.round_out ? TF_ROUND : 0) | (p.colsum ? TF_COLSUM : 0) | (p.mask_out ? TF_MASK_OUT : 0) |
            (p.mask_in ? TF_MASK_IN : 0) | (p.loss_partials ? TF_LOSS : 0) | (extra_flags & (TF_HALF | TF_SIG_HALF));
  t.act = p.act; t.M = p.M; t.N = p.N;
  g->tasks.push_back(t);
  g->uploaded = false;
  return (int)g->tasks.size() - 1;
}

// an elementwise task over row block m_blk (kind 0: latent forward, 1: latent backward; arguments: group_set_elem)
int group_add_elt_task(GroupPlan* g, int kind, int m_blk, int batch, int wait_ctr, int wait_cnt, int wait_val,
                       int wait2_ctr, int wait2_val, int signal_ctr, int wait2_cnt, int variant) {
  GTask t;
  memset(&t, 0, sizeof t);
  t.problem = (int)g->problems.size() > 0 ? g->sites.back().first_problem : 0;   // unused; keeps the site-relative rebase >= 0
  t.m_blk = m_blk; t.nkb = 0; t.kb0 = wait2_cnt;
  t.wait_ctr = wait_ctr; t.wait_cnt = wait_cnt; t.wait_val = wait_val;
  t.wait2_ctr = wait2_ctr; t.wait2_val = wait2_val; t.signal_ctr = signal_ctr;
  t.bn = 64;
  t.flags = kind == 0 ? TF_ELT_LATENT_FWD : kind == 1 ? TF_ELT_LATENT_BWD : TF_ELT_FINALIZE;
  t.act = variant;           // latent forward: 1 = the heads are split-K sums, add the layer's bias (LatentArgs::head_bias)
  t.M = batch;
  g->tasks.push_back(t);
  g->uploaded = false;
  return (int)g->tasks.size() - 1;
}

void group_set_elem(GroupPlan* g, const GElem& e) { g->elem = e; }

int group_num_tasks(const GroupPlan* g) { return (int)g->tasks.size(); }

void group_set_counters(GroupPlan* g, uint32_t* d_counters, int n, int half_off) {
  g->d_counters = d_counters; g->n_counters = n; g->half_off = half_off;
}
int group_problem_bn(const GroupPlan* g, int prob) { return g->problems[prob].BN; }
// a launch site must not mix 32- and 64-deep problems: a plan whose fused sites contain K-major operands with a pitch
// that is neither <= 32 nor a multiple of 32 floats (z / d heads at unusual latent widths) switches 64-deep k-blocks off
void group_set_deep_k(GroupPlan* g, bool allow) { g->allow_k64 = allow; }

bool group_upload(GroupPlan* g, char* err, int errlen) {
  if (g->uploaded) return true;
  if (g->d_tasks) { cudaFree(g->d_tasks); g->d_tasks = nullptr; }
  if (g->tasks.empty()) { g->uploaded = true; return true; }
  cudaError_t e = cudaMalloc(&g->d_tasks, g->tasks.size() * sizeof(GTask));
  if (e == cudaSuccess) e = cudaMemcpy(g->d_tasks, g->tasks.data(), g->tasks.size() * sizeof(GTask), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    static bool attr_done = false;
    if (!attr_done) {
      e = cudaFuncSetAttribute(gemm_group_kernel<kSiteProblemsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_group_kernel<kSiteProblemsLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_group_kernel<kSiteProblemsStep>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      attr_done = (e == cudaSuccess);
    }
  }
  if (e != cudaSuccess) { snprintf(err, errlen, "group upload failed: %s", cudaGetErrorString(e)); return false; }
  g->uploaded = true;
  return true;
}

namespace {
void launch_site(const GroupPlan* g, int site, uint32_t* queue, int reset_first, int reset_count, int dynamic_first,
                 unsigned long long* tl, cudaStream_t s, int advance = 0) {
  const GroupSite& st = g->sites[site];
  if (st.n_tasks <= 0) return;
  static const int env_mode = (getenv("VAEASSOC_EPI_BIAS_SHFL") ? 0 : 2) | (getenv("VAEASSOC_EPI_TMA_STORE") ? 4 : 0) |
                              (getenv("VAEASSOC_DEBUG_SKIP_MATH") ? 8 : 0) | (getenv("VAEASSOC_DEBUG_SKIP_STORE") ? 16 : 0) |
                              (getenv("VAEASSOC_EPI_STAGGER_NS") ? (std::max(0, std::min(4000, atoi(getenv("VAEASSOC_EPI_STAGGER_NS")))) << 8) : 0);   // bits 8..19
  int max_bn = 0;
  for (int i = 0; i < st.n_problems; ++i) max_bn = std::max(max_bn, g->problems[st.first_problem + i].BN);
  static const bool ring_fixed = getenv("VAEASSOC_RING_FIXED") != nullptr;
  int n_k64 = 0;
  for (int i = 0; i < st.n_problems; ++i) n_k64 += g->problems[st.first_problem + i].k64 ? 1 : 0;
  if (n_k64 != 0 && n_k64 != st.n_problems) {
    fprintf(stderr, "[vaeassoc] launch site mixes 32- and 64-deep k-blocks (%d of %d problems): plan bug\n", n_k64, st.n_problems);
    abort();
  }
  const int ring_class = n_k64 ? 3 : ring_fixed ? 0 : (max_bn <= 64 ? 2 : (max_bn <= 128 ? 1 : 0));
  const int mode = (dynamic_first ? 1 : 0) | env_mode | (advance ? 32 : 0) | (ring_class << 6) | (small_rows_enabled() ? (1 << 24) : 0);
  const int clusters = std::min(st.n_tasks, kNumSMs / 2);
  if (st.n_problems <= kSiteProblemsSmall) {
    GParams<kSiteProblemsSmall> prm;
    memcpy(prm.p, g->problems.data() + st.first_problem, (size_t)st.n_problems * sizeof(GProblem));
    gemm_group_kernel<kSiteProblemsSmall><<<2 * clusters, kThreads, SMEM_BYTES, s>>>(
        prm, g->elem, g->d_tasks + st.first_task, st.n_tasks, g->d_counters, queue, reset_first, reset_count, mode, tl, g->half_off);
  } else if (st.n_problems <= kSiteProblemsLarge) {
    GParams<kSiteProblemsLarge> prm;
    memcpy(prm.p, g->problems.data() + st.first_problem, (size_t)st.n_problems * sizeof(GProblem));
    gemm_group_kernel<kSiteProblemsLarge><<<2 * clusters, kThreads, SMEM_BYTES, s>>>(
        prm, g->elem, g->d_tasks + st.first_task, st.n_tasks, g->d_counters, queue, reset_first, reset_count, mode, tl, g->half_off);
  } else {
    GParams<kSiteProblemsStep> prm;
    memcpy(prm.p, g->problems.data() + st.first_problem, (size_t)st.n_problems * sizeof(GProblem));
    gemm_group_kernel<kSiteProblemsStep><<<2 * clusters, kThreads, SMEM_BYTES, s>>>(
        prm, g->elem, g->d_tasks + st.first_task, st.n_tasks, g->d_counters, queue, reset_first, reset_count, mode, tl, g->half_off);
  }
}
}  // namespace

// launches a site; the grid is one CTA pair per TPC at most.  `queue` = two zero-initialised words owned by this
// launch site (task-queue head, clusters-left count; the kernel rewinds them and the counters
// [reset_first, +reset_count) when its last cluster leaves)
void group_launch(const GroupPlan* g, int site, uint32_t* queue, int reset_first, int reset_count, int dynamic_first,
                  cudaStream_t s, int advance) {
  launch_site(g, site, queue, reset_first, reset_count, dynamic_first, nullptr, s, advance);
}

// debug only (VAEASSOC_TC_TIMELINE): one more run of the range with %globaltimer stamps per task; prints, relative to
// the first cluster's entry (ns): producer start, MMA start, MMA done, epilogue start, epilogue end, cluster
void group_debug_timeline(const GroupPlan* g, int site, uint32_t* queue, int reset_first, int reset_count, cudaStream_t s) {
  const int first = g->sites[site].first_task, count = g->sites[site].n_tasks;
  if (count <= 0) return;
  if (!kTimeline) { fprintf(stderr, "[group timeline] not compiled in: python -m vae_assoc_b200.build --timeline, VAEASSOC_LIB=.../libvaeassoc_tl.so\n"); return; }
  unsigned long long* dev = nullptr;
  if (cudaMalloc(&dev, (size_t)count * kTL * 8 + 192 * 8) != cudaSuccess) return;
  cudaMemsetAsync(dev, 0, (size_t)count * kTL * 8 + 192 * 8, s);
  const int clusters = std::min(count, kNumSMs / 2);
  launch_site(g, site, queue, reset_first, reset_count, 0, dev, s);
  cudaStreamSynchronize(s);
  std::vector<unsigned long long> h((size_t)count * kTL + 192);
  cudaMemcpy(h.data(), dev, (size_t)count * kTL * 8 + 192 * 8, cudaMemcpyDeviceToHost);
  cudaFree(dev);
  unsigned long long t0 = ~0ull, t1 = 0;
  for (int i = 0; i < count; ++i) {
    if (g->tasks[first + i].nkb == 0) continue;      // elementwise tasks carry no stamps
    t0 = std::min(t0, h[kTL * i + 6]); t1 = std::max(t1, h[kTL * i + 4]);
  }
  fprintf(stderr, "[group timeline] %d tasks on %d clusters, %.1f us from first entry to last epilogue end\n", count, clusters,
          (t1 - t0) * 1e-3);
  {
    const int nk = std::min(g->tasks[first].nkb, 64);
    fprintf(stderr, "  task 0 k-blocks (us): leader issue | peer issue | full seen by MMA\n   ");
    for (int i = 0; i < nk; ++i)
      fprintf(stderr, " %d: %.2f|%.2f|%.2f", i, (h[kTL * count + i] - t0) * 1e-3, (h[kTL * count + 128 + i] - t0) * 1e-3,
              (h[kTL * count + 64 + i] - t0) * 1e-3);
    fprintf(stderr, "\n");
  }
  const int show = getenv("VAEASSOC_TC_TIMELINE_ALL") ? count : std::min(count, 12);
  for (int k = 0; k < show; ++k) {
    const int i = (k < show / 2 || show == count) ? k : count - (show - k);
    const GTask& tk = g->tasks[first + i];
    if (tk.nkb == 0) {
      fprintf(stderr, "  elt  %4d flags %5d rb %2d cl %2llu | popped %6.2f inputs ready %6.2f work done %6.2f signalled %6.2f us\n", i, tk.flags,
              tk.m_blk, h[kTL * i + 5], (h[kTL * i + 3] - t0) * 1e-3, (h[kTL * i + 0] - t0) * 1e-3, (h[kTL * i + 8] - t0) * 1e-3,
              (h[kTL * i + 4] - t0) * 1e-3);
      continue;
    }
    fprintf(stderr, "  task %4d prob %2d (%2d,%2d) nkb %3d cl %2llu entry %6.2f | prod %6.2f mma %6.2f..%6.2f (%llu cyc) epi %6.2f..%6.2f us | ld0 %6.0f chunk0 %6.0f loop %6.2f | math %6.0f mask %6.0f wait %6.0f sts %6.0f\n", i,
            tk.problem, tk.m_blk, tk.n_blk, tk.nkb, h[kTL * i + 5], (h[kTL * i + 6] - t0) * 1e-3, (h[kTL * i + 0] - t0) * 1e-3,
            (h[kTL * i + 1] - t0) * 1e-3, (h[kTL * i + 2] - t0) * 1e-3, h[kTL * i + 7], (h[kTL * i + 3] - t0) * 1e-3, (h[kTL * i + 4] - t0) * 1e-3,
(double)(long long)(h[kTL * i + 9] - h[kTL * i + 15]), (double)(long long)(h[kTL * i + 10] - h[kTL * i + 15]), (h[kTL * i + 8] - t0) * 1e-3,
            (double)(long long)(h[kTL * i + 11] - h[kTL * i + 15]), (double)(long long)(h[kTL * i + 12] - h[kTL * i + 15]),
            (double)(long long)(h[kTL * i + 13] - h[kTL * i + 15]), (double)(long long)(h[kTL * i + 14] - h[kTL * i + 15]));
  }
}

}  // namespace vaeassoc
