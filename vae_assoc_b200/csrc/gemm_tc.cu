// tcgen05 / TMA dense-layer GEMMs for sm_100a -- the tensor-core path of the assoc-VAE train step.
//
//   kind NN  C[M,N]  = act(A[M,K] . B[K,N] + bias)          forward   (vae_assoc.py:187-188,203-204,282-283,295-303)
//   kind NT  C[M,N]  = (A[M,K] . B[N,K]^T) (*) act'(aux)    dgrad     (autodiff of the above, :373-374)
//   kind TN  C[M,N] += A[K,M]^T . B[K,N]                    wgrad     (K = batch, split across CTAs, fp32 RED)
//
// Operands are fp32 in HBM, already rounded to tf32 by their producers; `tcgen05.mma.kind::tf32` reads them from
// shared memory (128-byte swizzle, written by TMA) and accumulates fp32 in TMEM.
//
// CTA = 192 threads: warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator + MMA issuer (one lane),
// warps 2..5 = epilogue (tcgen05.ld -> bias/activation/act-grad -> global).  One 128 x BN output tile per CTA,
// BK = 32 fp32 (= one 128-byte swizzle row) per pipeline stage; two CTAs are co-resident per SM so that one CTA's
// epilogue overlaps the other's main loop.
//
// Shared-memory operand layouts (both are canonical UMMA layouts, descriptors below):
//   K-major  operand (contraction contiguous in HBM): tile [R rows][32 k] -> R rows of 128 B, one TMA box {32, R}.
//            UMMA desc: SWIZZLE_128B, SBO = 1024 B (8 rows); the four K=8 MMAs of a stage advance the start
//            address by 32 B.
//   MN-major operand (M/N index contiguous in HBM):   tile [R/32 chunks][32 k rows][32 mn] -> R/32 TMA boxes
//            {32, 32} of 4 KB.  For 32-bit (tf32) MN-major operands the only UMMA layout is SWIZZLE_128B_BASE32B
//            (32-byte chunks XOR-ed with the row index mod 4; TMA mode SWIZZLE_128B_ATOM_32B): LBO = 4096 B (next
//            32-wide MN chunk), SBO = 512 B (next 4 k rows); the four MMAs of a stage advance the start by 1024 B.
// TMA zero-fills out-of-range rows/columns, so M, N, K need no padding (only 16-byte row pitches).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace vaeassoc {

namespace {

constexpr int BM = 128;          // UMMA M (one TMEM lane per output row)
constexpr int BK = 32;           // fp32 elements per stage = one 128-byte swizzle row
constexpr int UMMA_K = 8;        // tf32: 32 bytes per instruction
constexpr int kThreads = 192;
constexpr int kStages128 = 3;    // BN = 128: 3 x 32 KB per CTA, 2 CTAs / SM

enum Kind : int { KIND_NN = 0, KIND_NT = 1, KIND_TN = 2 };

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
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
// K-major tile: rows of 128 B, 8-row groups 1024 B apart.  MN-major tile: 32-wide chunks 4096 B apart, 4-row groups 512 B
__device__ __forceinline__ uint64_t desc_k_major(uint32_t addr) { return make_desc(addr, 16, 1024, 2); }
__device__ __forceinline__ uint64_t desc_mn_major(uint32_t addr) { return make_desc(addr, BK * 128, 512, 1); }

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format [4,6)=1 (F32), a/b_format [7,10),[10,13)=2 (TF32),
// a_major bit 15, b_major bit 16 (1 = MN-major), n_dim [17,23) = N>>3, m_dim [24,29) = M>>4
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct TcArgs {
  int M, N, K;               // output rows, output cols, contraction length
  float* C; int64_t ldc;
  const float* bias;
  const float* aux; int64_t ldaux;
  int act, round_out;
  int kblocks_per_split;     // BK-blocks of the contraction handled by one CTA (blockIdx.z)
  unsigned long long* timeline;   // debug only (VAEASSOC_TC_TIMELINE): 8 stamps per CTA, else null
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned smid() {
  unsigned r;
  asm volatile("mov.u32 %0, %smid;" : "=r"(r));
  return r;
}

template <int KIND, int BN>
struct Cfg {
  static constexpr bool A_MN = (KIND == KIND_TN);
  static constexpr bool B_MN = (KIND != KIND_NT);
  static constexpr int STAGES = kStages128;
  static constexpr int A_BYTES = BM * BK * 4;
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int KIND, int BN>
__global__ void __launch_bounds__(kThreads, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TcArgs g) {
  using C = Cfg<KIND, BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + C::STAGES * C::STAGE_BYTES;
  // barriers: full[s], empty[s], tmem_full ; then the TMEM base address slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * C::STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * C::STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int total_kb = (g.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * g.kblocks_per_split;
  const int kb1 = min(total_kb, kb0 + g.kblocks_per_split);
  const int nkb = kb1 - kb0;
  unsigned long long* tl = g.timeline
      ? g.timeline + 8ull * (blockIdx.x + gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z)) : nullptr;
  if (tl && threadIdx.x == 0) { tl[0] = gtimer(); tl[1] = clock64(); tl[7] = smid(); }

  if (threadIdx.x == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (tl && threadIdx.x == 0) tl[2] = clock64();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % C::STAGES;
        mbar_wait(empty_bar(s), ((i / C::STAGES) & 1) ^ 1);
        const uint32_t sa = base + s * C::STAGE_BYTES, sb = sa + C::A_BYTES;
        const int k0 = (kb0 + i) * BK;
        mbar_arrive_expect_tx(full_bar(s), C::STAGE_BYTES);
        if (C::A_MN) {
#pragma unroll
          for (int j = 0; j < BM / 32; ++j) tma_load_2d(sa + j * (BK * 128), &map_a, full_bar(s), m0 + 32 * j, k0);
        } else {
          tma_load_2d(sa, &map_a, full_bar(s), k0, m0);
        }
        if (C::B_MN) {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(sb + j * (BK * 128), &map_b, full_bar(s), n0 + 32 * j, k0);
        } else {
          tma_load_2d(sb, &map_b, full_bar(s), k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN, C::A_MN, C::B_MN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % C::STAGES;
        mbar_wait(full_bar(s), (i / C::STAGES) & 1);
        tc_fence_after();
        if (tl && i == 0) tl[3] = clock64();
        const uint32_t sa = base + s * C::STAGE_BYTES, sb = sa + C::A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t da = C::A_MN ? desc_mn_major(sa + k * 1024) : desc_k_major(sa + k * 32);
          const uint64_t db = C::B_MN ? desc_mn_major(sb + k * 1024) : desc_k_major(sb + k * 32);
          umma_tf32(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));          // frees the smem slot once these MMAs have read it
      }
      umma_commit(tmem_full_bar);           // accumulator complete
      if (tl) tl[4] = clock64();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // TMEM -> registers (thread = output row, 32 columns) -> shared-memory transpose -> global, so that every
    // global access of the epilogue (bias, stored activation for act', C) is a coalesced 128-byte row segment.
    // The staging tile aliases pipeline stage 0: when tmem_full fires every TMA load has landed and every MMA has
    // consumed its operands, so the stage buffers are dead.
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    float* stg = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw))) + q * (32 * 33);
    const int row0 = m0 + q * 32;
    const int nrows = min(32, g.M - row0);  // warp-uniform
    const bool use_aux = (KIND != KIND_TN) && g.aux != nullptr;   // NN + aux: v * act'(aux) instead of act(v)
    // act'(stored activation) does not depend on the accumulator: its 32 x 128-byte row segments per chunk are
    // requested up front (32 independent loads in flight per lane), the first chunk while the main loop still runs
    float auxv[32];
    auto load_aux = [&](int c) {
      const int col = n0 + c * 32 + lane;
#pragma unroll
      for (int r = 0; r < 32; ++r)
        auxv[r] = (use_aux && r < nrows && col < g.N) ? g.aux[(int64_t)(row0 + r) * g.ldaux + col] : 1.0f;
    };
    if (use_aux) load_aux(0);
    if (nkb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    if (tl && warp == 2 && lane == 0) tl[5] = clock64();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      if (nkb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      const int col = n0 + c * 32 + lane;
      if (n0 + c * 32 >= g.N || nrows <= 0) continue;          // warp-uniform
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);   // bank (lane + j) % 32: conflict-free
      __syncwarp();
      const bool col_ok = col < g.N;
      const float bias_v = (KIND == KIND_NN && g.bias && col_ok) ? __ldg(g.bias + col) : 0.f;
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        if (r < nrows && col_ok) {
          float x = stg[r * 33 + lane];
          float* dst = g.C + (int64_t)(row0 + r) * g.ldc + col;
          if (KIND == KIND_TN) {
            atomicAdd(dst, x);                                  // split-K reduction: one 128-byte RED per warp
          } else {
            if (KIND == KIND_NN) x += bias_v;
            if (use_aux) x *= act_grad_from_output(g.act, auxv[r]);
            else if (KIND == KIND_NN) x = apply_act(g.act, x);
            if (g.round_out) x = round_tf32(x);
            *dst = x;
          }
        }
      }
      __syncwarp();
      if (use_aux && c + 1 < BN / 32) load_aux(c + 1);
    }
  }
  if (tl && warp == 2 && lane == 0) tl[6] = clock64();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
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

// 2-D fp32 tensor map: dim0 (contiguous) x dim1 with row pitch `ld` floats, box {32, box_rows}, 128-byte swizzle
bool make_map(CUtensorMap* map, const float* ptr, int64_t dim0, int64_t dim1, int64_t ld, int box_rows, bool mn_major,
              char* err, int errlen) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { snprintf(err, errlen, "cuTensorMapEncodeTiled entry point not available"); return false; }
  cuuint64_t dims[2] = {(cuuint64_t)dim0, (cuuint64_t)dim1};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
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
  CUtensorMap map_a, map_b;
  TcArgs args;
  dim3 grid;
};

bool tc_supported(int kind, const GemmArgs& a) {
  if (!a.A || !a.B || !a.C) return false;
  if ((reinterpret_cast<uintptr_t>(a.A) & 15) || (reinterpret_cast<uintptr_t>(a.B) & 15)) return false;
  if ((a.lda & 3) || (a.ldb & 3)) return false;
  // worth a 128 x 128 tensor-core tile: skinny shapes (n_z-wide heads, K = n_z) stay on the SIMT kernels
  if (a.M < 32 || a.N < 32 || a.K < 32) return false;
  (void)kind;
  return true;
}

TcPlan* tc_plan_create(int kind, const GemmArgs& a, char* err, int errlen) {
  TcPlan* p = new TcPlan();
  p->kind = kind;
  constexpr int BN = 128;
  bool ok = true;
  switch (kind) {
    case KIND_NN:   // A [M,K] K-major ; B [K,N] MN-major
      ok = make_map(&p->map_a, a.A, a.K, a.M, a.lda, BM, false, err, errlen) &&
           make_map(&p->map_b, a.B, a.N, a.K, a.ldb, BK, true, err, errlen);
      break;
    case KIND_NT:   // A [M,K] K-major ; B [N,K] K-major
      ok = make_map(&p->map_a, a.A, a.K, a.M, a.lda, BM, false, err, errlen) &&
           make_map(&p->map_b, a.B, a.K, a.N, a.ldb, BN, false, err, errlen);
      break;
    default:        // A [K,M] MN-major ; B [K,N] MN-major
      ok = make_map(&p->map_a, a.A, a.M, a.K, a.lda, BK, true, err, errlen) &&
           make_map(&p->map_b, a.B, a.N, a.K, a.ldb, BK, true, err, errlen);
      break;
  }
  if (!ok) { delete p; return nullptr; }
  TcArgs& t = p->args;
  t.M = a.M; t.N = a.N; t.K = a.K; t.C = a.C; t.ldc = a.ldc; t.bias = a.bias; t.aux = a.aux; t.ldaux = a.ldaux;
  t.act = a.act; t.round_out = a.round_out; t.timeline = nullptr;
  const int tiles_m = (a.M + BM - 1) / BM, tiles_n = (a.N + BN - 1) / BN;
  const int total_kb = (a.K + BK - 1) / BK;
  int splits = 1;
  if (kind == KIND_TN) {
    // split the batch contraction so that ~2 CTAs per SM are busy; each split keeps >= 4 k-blocks
    const int want = (2 * kNumSMs + tiles_m * tiles_n - 1) / (tiles_m * tiles_n);
    splits = std::max(1, std::min(want, total_kb / 4));
    if (a.splitk > 1) splits = std::min(a.splitk, total_kb);
  }
  t.kblocks_per_split = (total_kb + splits - 1) / splits;
  splits = (total_kb + t.kblocks_per_split - 1) / t.kblocks_per_split;
  p->grid = dim3(tiles_n, tiles_m, splits);
  static bool attr_done[3] = {false, false, false};
  if (!attr_done[kind]) {
    cudaError_t e = cudaSuccess;
    switch (kind) {
      case KIND_NN: e = cudaFuncSetAttribute(gemm_tc_kernel<KIND_NN, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<KIND_NN, BN>::SMEM); break;
      case KIND_NT: e = cudaFuncSetAttribute(gemm_tc_kernel<KIND_NT, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<KIND_NT, BN>::SMEM); break;
      default: e = cudaFuncSetAttribute(gemm_tc_kernel<KIND_TN, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<KIND_TN, BN>::SMEM); break;
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
  auto med = [&](int a, int b) {
    std::vector<long long> v;
    for (size_t i = 0; i < n; ++i) v.push_back((long long)(h[8 * i + b] - h[8 * i + a]));
    std::sort(v.begin(), v.end());
    fprintf(stderr, " [%d->%d] min %lld med %lld max %lld |", a, b, v[0], v[n / 2], v[n - 1]);
  };
  unsigned long long g0 = ~0ull, g1 = 0;
  for (size_t i = 0; i < n; ++i) { g0 = std::min(g0, h[8 * i]); g1 = std::max(g1, h[8 * i]); }
  fprintf(stderr, "[tc timeline] kind %d grid (%u,%u,%u) start spread %llu ns; cycles:", p->kind, p->grid.x, p->grid.y, p->grid.z, g1 - g0);
  med(1, 2); med(2, 3); med(3, 4); med(4, 5); med(5, 6); med(1, 6);
  fprintf(stderr, "\n");
}

void launch_gemm_tc(const TcPlan* p, cudaStream_t s) {
  constexpr int BN = 128;
  switch (p->kind) {
    case KIND_NN: gemm_tc_kernel<KIND_NN, BN><<<p->grid, kThreads, Cfg<KIND_NN, BN>::SMEM, s>>>(p->map_a, p->map_b, p->args); break;
    case KIND_NT: gemm_tc_kernel<KIND_NT, BN><<<p->grid, kThreads, Cfg<KIND_NT, BN>::SMEM, s>>>(p->map_a, p->map_b, p->args); break;
    default: gemm_tc_kernel<KIND_TN, BN><<<p->grid, kThreads, Cfg<KIND_TN, BN>::SMEM, s>>>(p->map_a, p->map_b, p->args); break;
  }
}

}  // namespace vaeassoc
