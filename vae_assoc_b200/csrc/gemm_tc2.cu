// tcgen05 / TMA dense-layer GEMMs for sm_100a, CTA-pair version -- the tensor-core path of the assoc-VAE train step.
//
//   kind NN  C[M,N]  = act(A[M,K] . B[K,N] + bias)          forward   (vae_assoc.py:187-188,203-204,282-283,295-303)
//   kind NT  C[M,N]  = (A[M,K] . B[N,K]^T) (*) act'(aux)    dgrad     (autodiff of the above, :373-374)
//   kind TN  C[M,N] += A[K,M]^T . B[K,N]                    wgrad     (K = batch, split across clusters, TMA reduce-add)
//
// Operands are fp32 in HBM, already rounded to tf32 by their producers; `tcgen05.mma.kind::tf32` reads them from
// shared memory (128-byte swizzle, written by TMA) and accumulates fp32 in TMEM.
//
// One cluster of two CTAs (one TPC) owns a 256 x BN output tile, BN in {64,128,192,256}: `tcgen05.mma.cta_group::2`
// with UMMA M = 256.  Each CTA stages its own 128 rows of A and its own half (BN/2 columns) of B, so a k-block of
// 32 costs 16 KB + BN*64 B of L2->SM traffic per CTA for 128 x BN x 32 MACs -- half the bytes per FLOP of a
// single-CTA 128 x 128 tile, which is what bounded the first version of this kernel (L2 fabric, not the tensor pipe).
// The leader CTA (cluster rank 0) issues every MMA; both CTAs' TMA loads signal the leader's `full` barrier; the
// leader's `tcgen05.commit` multicasts to both CTAs' `empty` / `tmem_full` barriers.
//
// CTA = 320 threads: warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (leader: one lane),
// warps 2..9 = epilogue (two per TMEM lane quarter, alternating 32-column chunks).  Epilogue per warp and chunk:
// tcgen05.ld (thread = output row, 32 columns) -> bias / activation / act' (aux tile fetched by TMA into swizzled
// smem, double buffered) -> swizzled smem staging (in the pipeline stages, dead once the accumulator is complete,
// so no store ever waits for a buffer) -> one TMA store (NN, NT) or TMA reduce-add (TN, the split-K reduction) of
// the 32 x 32 box.  TMA clips the boxes at the M / N edges (in 16-byte units: columns N .. roundup4(N) receive
// zeros) and zero-fills loads past M / N / K, so no dimension needs padding (only 16-byte row pitches).
//
// Shared-memory operand layouts (canonical UMMA layouts):
//   K-major  operand (contraction contiguous in HBM): tile [R rows][32 k] -> R rows of 128 B, one TMA box {32, R}.
//            UMMA desc: SWIZZLE_128B, SBO = 1024 B (8 rows); the four K=8 MMAs of a stage advance the start by 32 B.
//   MN-major operand (M/N index contiguous in HBM):   tile [R/32 chunks][32 k rows][32 mn] -> R/32 TMA boxes
//            {32, 32} of 4 KB.  For 32-bit (tf32) MN-major operands the only UMMA layout is SWIZZLE_128B_BASE32B
//            (TMA mode SWIZZLE_128B_ATOM_32B): LBO = 4096 B (next 32-wide MN chunk), SBO = 512 B (next 4 k rows);
//            the four MMAs of a stage advance the start by 1024 B.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace vaeassoc {

namespace {

constexpr int BM_CTA = 128;      // rows of the tile per CTA; UMMA M = 256 over the pair
constexpr int BM = 2 * BM_CTA;
constexpr int BK = 32;           // fp32 elements per stage = one 128-byte swizzle row
constexpr int UMMA_K = 8;        // tf32: 32 bytes per instruction
constexpr int kEpiWarps = 8;     // two warps per TMEM lane quarter, each takes every other 32-column chunk
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kStages = 4;
constexpr int A_BYTES = BM_CTA * BK * 4;          // 16 KB
constexpr int B_BYTES_MAX = 128 * BK * 4;         // BN/2 <= 128 columns
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;
constexpr int CHUNK_BYTES = 32 * 32 * 4;          // one 32 x 32 fp32 epilogue box
constexpr int EPI_WARP_BYTES = 2 * CHUNK_BYTES;   // 2 aux buffers per epilogue warp (output staging reuses the stages)
constexpr int kTmemCols = 256;
constexpr int SMEM_BYTES = kStages * STAGE_BYTES + kEpiWarps * EPI_WARP_BYTES + 256 * 4 /*bias*/ + 256 /*barriers*/ + 1024 /*align*/;
static_assert(kStages * STAGE_BYTES >= 4 * 8 * CHUNK_BYTES, "the dead pipeline stages must hold the CTA's whole 128 x 256 output");

enum Kind : int { KIND_NN = 0, KIND_NT = 1, KIND_TN = 2 };

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
// bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    if (spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// operand tile load of the CTA pair: data lands in this CTA's shared memory, the bytes are counted on `bar`, a
// shared::cluster address (the leader CTA's `full` barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
// CTA-local tile load (epilogue aux tiles)
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
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
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
__device__ __forceinline__ uint64_t desc_mn_major(uint32_t addr) { return make_desc(addr, BK * 128, 512, 1); }

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format [4,6)=1 (F32), a/b_format [7,10),[10,13)=2 (TF32),
// a_major bit 15, b_major bit 16 (1 = MN-major), n_dim [17,23) = N>>3, m_dim [24,29) = M>>4
__device__ __forceinline__ uint32_t make_idesc(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct Tc2Args {
  int M, N, K;               // output rows, output cols, contraction length
  int BN;                    // tile width: 64, 128, 192 or 256
  const float* bias;
  int has_aux;               // epilogue multiplies by act'(aux tile) instead of applying act
  int act, round_out;
  int kblocks_per_split;     // BK-blocks of the contraction handled by one cluster (blockIdx.z)
  unsigned long long* timeline;   // debug only (VAEASSOC_TC_TIMELINE): 8 stamps per CTA, else null
};

// byte offset of 16-byte chunk j of row r inside a 32 x 32 fp32 box written / read by TMA with SWIZZLE_128B
__device__ __forceinline__ uint32_t swz(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }

template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_aux, Tc2Args g) {
  constexpr bool A_MN = (KIND == KIND_TN);
  constexpr bool B_MN = (KIND != KIND_NT);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t epi_base = base + kStages * STAGE_BYTES;
  const uint32_t bias_base = epi_base + kEpiWarps * EPI_WARP_BYTES;
  const uint32_t bar_base = bias_base + 256 * 4;
  // barriers: full[s] (used in the leader CTA only), empty[s], tmem_full, aux_full[epilogue warp][2]; then the TMEM slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kStages);
  auto aux_bar = [&](int e, int b) { return bar_base + 8u * (2 * kStages + 1 + 2 * e + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 1 + 2 * kEpiWarps);
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));   // generic pointer to `base`
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - base));
  float* bias_s = reinterpret_cast<float*>(smem_gen + (bias_base - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int BN = g.BN, BNH = BN >> 1;
  const int m0 = (int)(blockIdx.x >> 1) * BM + (int)rank * BM_CTA;   // this CTA's rows of A and of the output
  const int n0 = (int)blockIdx.y * BN;                               // the tile's first column
  const int nb0 = n0 + (int)rank * BNH;                              // this CTA's slice of B
  const int total_kb = (g.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * g.kblocks_per_split;
  const int nkb = min(total_kb, kb0 + g.kblocks_per_split) - kb0;    // > 0 by construction of the grid
  unsigned long long* tl = g.timeline
      ? g.timeline + 8ull * (blockIdx.x + gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z)) : nullptr;
  if (tl && threadIdx.x == 0) { tl[0] = gtimer(); tl[1] = clock64(); }

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    prefetch_tensormap(&map_c);
    if (g.has_aux) prefetch_tensormap(&map_aux);
    for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    for (int e = 0; e < kEpiWarps; ++e) { mbar_init(aux_bar(e, 0), 1); mbar_init(aux_bar(e, 1), 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, kTmemCols);
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < BN; i += 32 * kEpiWarps)
      bias_s[i] = (KIND == KIND_NN && g.bias != nullptr && n0 + i < g.N) ? __ldg(g.bias + n0 + i) : 0.0f;
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (tl && threadIdx.x == 0) tl[2] = clock64();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      const uint32_t b_bytes = (uint32_t)BNH * BK * 4;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % kStages;
        mbar_wait(empty_bar(s), ((i / kStages) & 1) ^ 1);
        const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_BYTES;
        const uint32_t full_leader = mapa(full_bar(s), 0);
        const int k0 = (kb0 + i) * BK;
        if (rank == 0) mbar_arrive_expect_tx(full_bar(s), 2u * (A_BYTES + b_bytes));
        if (A_MN) {
#pragma unroll
          for (int j = 0; j < BM_CTA / 32; ++j) tma_load_2d_pair(sa + j * (BK * 128), &map_a, full_leader, m0 + 32 * j, k0);
        } else {
          tma_load_2d_pair(sa, &map_a, full_leader, k0, m0);
        }
        if (B_MN) {
          for (int j = 0; j < BNH / 32; ++j) tma_load_2d_pair(sb + j * (BK * 128), &map_b, full_leader, nb0 + 32 * j, k0);
        } else {
          tma_load_2d_pair(sb, &map_b, full_leader, k0, nb0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc(BM, BN, A_MN, B_MN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % kStages;
        mbar_wait(full_bar(s), (i / kStages) & 1);
        tc_fence_after();
        if (tl && i == 0) tl[3] = clock64();
        const uint32_t sa = base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t da = A_MN ? desc_mn_major(sa + k * 1024) : desc_k_major(sa + k * 32);
          const uint64_t db = B_MN ? desc_mn_major(sb + k * 1024) : desc_k_major(sb + k * 32);
          umma_tf32_pair(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit_pair(empty_bar(s));     // frees the smem slot in both CTAs once these MMAs have read it
      }
      umma_commit_pair(tmem_full_bar);      // accumulator complete (both CTAs)
      if (tl) tl[4] = clock64();
    }
  } else {
    // ===================== epilogue (warps 2..9 of both CTAs) =====================
    const int e = warp - 2;                 // epilogue warp index
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = e >> 2;                // this warp takes chunks half, half + 2, ...
    const int row0 = m0 + q * 32;           // first output row of this warp
    const uint32_t aux_buf = epi_base + e * EPI_WARP_BYTES;        // 2 x 4 KB aux tiles
    const int nchunks = (row0 < g.M) ? min(BN / 32, (g.N - n0 + 31) / 32) : 0;   // warp-uniform
    const bool use_aux = (KIND != KIND_TN) && g.has_aux;
    if (use_aux && half < nchunks && lane == 0) {
      mbar_arrive_expect_tx(aux_bar(e, 0), CHUNK_BYTES);
      tma_load_2d(aux_buf, &map_aux, aux_bar(e, 0), n0 + half * 32, row0);
    }
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    if (tl && warp == 2 && lane == 0) tl[5] = clock64();
#pragma unroll 1
    for (int c = half, it = 0; c < nchunks; c += 2, ++it) {
      const int b = it & 1;
      if (use_aux && c + 2 < nchunks && lane == 0) {          // prefetch the next aux tile (its buffer was last read at it-1)
        mbar_arrive_expect_tx(aux_bar(e, b ^ 1), CHUNK_BYTES);
        tma_load_2d(aux_buf + (b ^ 1) * CHUNK_BYTES, &map_aux, aux_bar(e, b ^ 1), n0 + (c + 2) * 32, row0);
      }
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      float x[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
      if (KIND == KIND_NN && !use_aux) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 bv = lds128(bias_base + (uint32_t)(c * 32 + j * 4) * 4);   // broadcast read
          x[4 * j] += bv.x; x[4 * j + 1] += bv.y; x[4 * j + 2] += bv.z; x[4 * j + 3] += bv.w;
        }
        switch (g.act) {
          case ACT_RELU:
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.0f);
            break;
          case ACT_SOFTPLUS:
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.0f) + __logf(1.0f + __expf(-fabsf(x[j])));
            break;
          case ACT_SIGMOID:
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __fdividef(1.0f, 1.0f + __expf(-x[j]));
            break;
          default: break;
        }
      }
      if (use_aux) {
        mbar_wait(aux_bar(e, b), (it >> 1) & 1);
        const uint32_t ab = aux_buf + b * CHUNK_BYTES;
        float h[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 hv = lds128(ab + swz(lane, j));
          h[4 * j] = hv.x; h[4 * j + 1] = hv.y; h[4 * j + 2] = hv.z; h[4 * j + 3] = hv.w;
        }
        switch (g.act) {
          case ACT_RELU:
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = h[j] > 0.0f ? x[j] : 0.0f;
            break;
          case ACT_SOFTPLUS:
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] *= 1.0f - __expf(-h[j]);
            break;
          case ACT_SIGMOID:
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] *= h[j] * (1.0f - h[j]);
            break;
          default: break;
        }
      }
      if (g.round_out) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = round_tf32(x[j]);
      }
      if (n0 + c * 32 + 32 > g.N) {        // TMA clips stores in 16-byte units: the pad columns N..roundup4(N) get zeros
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = (n0 + c * 32 + j < g.N) ? x[j] : 0.0f;
      }
      // staging: the pipeline stages are dead (every TMA load has landed, every MMA has read its operands)
      const uint32_t ob = base + (uint32_t)(q * 8 + c) * CHUNK_BYTES;
#pragma unroll
      for (int j = 0; j < 8; ++j) sts128(ob + swz(lane, j), x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        if (KIND == KIND_TN) tma_reduce_add_2d(&map_c, ob, n0 + c * 32, row0);
        else tma_store_2d(&map_c, ob, n0 + c * 32, row0);
        bulk_commit();
      }
    }
    if (tl && warp == 2 && lane == 0) tl[7] = clock64();
    if (lane == 0) bulk_wait_read<0>();     // shared memory must outlive the reads of the bulk stores
    if (tl && warp == 2 && lane == 0) tl[6] = clock64();
  }
  tc_fence_before();
  cluster_sync_all();     // the peer's smem / barriers stay valid until every MMA and every TMA of the pair is done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
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

}  // namespace

struct TcPlan {
  int kind = 0;
  CUtensorMap map_a, map_b, map_c, map_aux;
  Tc2Args args;
  dim3 grid;
};

bool tc_supported(int kind, const GemmArgs& a) {
  if (!a.A || !a.B || !a.C) return false;
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15) ||
      (reinterpret_cast<uintptr_t>(a.C) & 15))
    return false;
  if ((a.lda & 3) || (a.ldb & 3) || (a.ldc & 3)) return false;
  if (a.aux && ((reinterpret_cast<uintptr_t>(a.aux) & 15) || (a.ldaux & 3))) return false;
  // every contraction whose batch extent fills a tile row block runs here, including the n_z-wide ones (heads,
  // decoder input layer): those are HBM-bound and the TMA pipeline streams the one large operand exactly once;
  // out-of-range rows / columns of the narrow operand are zero-filled by TMA at no HBM cost
  const int batch = (kind == KIND_TN) ? a.K : a.M;
  return batch >= 32;
}

TcPlan* tc_plan_create(int kind, const GemmArgs& a, char* err, int errlen) {
  TcPlan* p = new TcPlan();
  p->kind = kind;
  const int tiles_n = (a.N + 255) / 256;
  const int BN = std::min(256, (((a.N + tiles_n - 1) / tiles_n) + 63) / 64 * 64);
  bool ok = true;
  switch (kind) {
    case KIND_NN:   // A [M,K] K-major ; B [K,N] MN-major
      ok = make_map(&p->map_a, a.A, a.K, a.M, a.lda, BM_CTA, false, err, errlen) &&
           make_map(&p->map_b, a.B, a.N, a.K, a.ldb, BK, true, err, errlen);
      break;
    case KIND_NT:   // A [M,K] K-major ; B [N,K] K-major
      ok = make_map(&p->map_a, a.A, a.K, a.M, a.lda, BM_CTA, false, err, errlen) &&
           make_map(&p->map_b, a.B, a.K, a.N, a.ldb, BN / 2, false, err, errlen);
      break;
    default:        // A [K,M] MN-major ; B [K,N] MN-major
      ok = make_map(&p->map_a, a.A, a.M, a.K, a.lda, BK, true, err, errlen) &&
           make_map(&p->map_b, a.B, a.N, a.K, a.ldb, BK, true, err, errlen);
      break;
  }
  ok = ok && make_map(&p->map_c, a.C, a.N, a.M, a.ldc, 32, false, err, errlen);
  const bool has_aux = kind != KIND_TN && a.aux != nullptr;
  if (ok && has_aux) ok = make_map(&p->map_aux, a.aux, a.N, a.M, a.ldaux, 32, false, err, errlen);
  else if (ok) p->map_aux = p->map_c;
  if (!ok) { delete p; return nullptr; }
  Tc2Args& t = p->args;
  t.M = a.M; t.N = a.N; t.K = a.K; t.BN = BN; t.bias = a.bias; t.has_aux = has_aux ? 1 : 0;
  t.act = a.act; t.round_out = a.round_out; t.timeline = nullptr;
  const int tiles_m = (a.M + BM - 1) / BM, tiles_nn = (a.N + BN - 1) / BN;
  const int total_kb = (a.K + BK - 1) / BK;
  int splits = 1;
  if (kind == KIND_TN) {
    // split the batch contraction so that every SM pair has a cluster; each split keeps >= 4 k-blocks
    const int pairs = kNumSMs / 2;
    const int want = pairs / (tiles_m * tiles_nn);      // floor: one wave of clusters
    splits = std::max(1, std::min(want, total_kb / 4));
    if (a.splitk > 1) splits = std::min(a.splitk, total_kb);
  }
  t.kblocks_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + t.kblocks_per_split - 1) / t.kblocks_per_split;
  p->grid = dim3(2 * tiles_m, tiles_nn, splits);
  static bool attr_done[3] = {false, false, false};
  if (!attr_done[kind]) {
    cudaError_t e = cudaSuccess;
    switch (kind) {
      case KIND_NN: e = cudaFuncSetAttribute(gemm_tc2_kernel<KIND_NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); break;
      case KIND_NT: e = cudaFuncSetAttribute(gemm_tc2_kernel<KIND_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); break;
      default: e = cudaFuncSetAttribute(gemm_tc2_kernel<KIND_TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES); break;
    }
    if (e != cudaSuccess) {
      snprintf(err, errlen, "cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
      delete p;
      return nullptr;
    }
    attr_done[kind] = true;
  }
  return p;
}

void tc_plan_destroy(TcPlan* p) { delete p; }

void launch_gemm_tc(const TcPlan* p, cudaStream_t s) {
  switch (p->kind) {
    case KIND_NN: gemm_tc2_kernel<KIND_NN><<<p->grid, kThreads, SMEM_BYTES, s>>>(p->map_a, p->map_b, p->map_c, p->map_aux, p->args); break;
    case KIND_NT: gemm_tc2_kernel<KIND_NT><<<p->grid, kThreads, SMEM_BYTES, s>>>(p->map_a, p->map_b, p->map_c, p->map_aux, p->args); break;
    default: gemm_tc2_kernel<KIND_TN><<<p->grid, kThreads, SMEM_BYTES, s>>>(p->map_a, p->map_b, p->map_c, p->map_aux, p->args); break;
  }
}

// debug only: run the plan once with per-CTA time stamps and print the phase medians (cycles) to stderr
void tc_debug_timeline(TcPlan* p, cudaStream_t s) {
  const size_t n = (size_t)p->grid.x * p->grid.y * p->grid.z;
  unsigned long long* dev = nullptr;
  if (cudaMalloc(&dev, n * 64) != cudaSuccess) return;
  cudaMemsetAsync(dev, 0, n * 64, s);
  p->args.timeline = dev;
  launch_gemm_tc(p, s);
  p->args.timeline = nullptr;
  std::vector<unsigned long long> h(n * 8);
  cudaStreamSynchronize(s);
  cudaMemcpy(h.data(), dev, n * 64, cudaMemcpyDeviceToHost);
  cudaFree(dev);
  auto med = [&](int a, int b, size_t first, size_t step) {
    std::vector<long long> v;
    for (size_t i = first; i < n; i += step) v.push_back((long long)(h[8 * i + b] - h[8 * i + a]));
    std::sort(v.begin(), v.end());
    fprintf(stderr, " [%d->%d] min %lld med %lld max %lld |", a, b, v[0], v[v.size() / 2], v[v.size() - 1]);
  };
  unsigned long long g0 = ~0ull, g1 = 0;
  for (size_t i = 0; i < n; ++i) { g0 = std::min(g0, h[8 * i]); g1 = std::max(g1, h[8 * i]); }
  fprintf(stderr, "[tc timeline] kind %d grid (%u,%u,%u) BN %d start spread %llu ns; leader cycles:", p->kind, p->grid.x,
          p->grid.y, p->grid.z, p->args.BN, g1 - g0);
  // grid.x is even and the leader CTA of a cluster has an even linear index
  med(1, 2, 0, 2); med(2, 3, 0, 2); med(3, 4, 0, 2); med(4, 5, 0, 2); med(5, 7, 0, 1); med(7, 6, 0, 1); med(1, 6, 0, 1);
  fprintf(stderr, "\n");
}

}  // namespace vaeassoc
