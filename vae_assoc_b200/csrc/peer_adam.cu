// Data-parallel parameter update over NVLink / NVSwitch peer memory: ONE kernel per step that
//   (1) reduce-scatters the flat gradient buffers  -- rank r sums, in rank order, shard r of every rank's `g`, read
//       straight from the peers' HBM through the NVLink peer mapping (cudaIpc), 128-bit system-scope loads;
//   (2) applies TensorFlow-formula Adam (vae_assoc.py:373-374) to its shard only (m, v are sharded: 1/world of the work);
//   (3) all-gathers the result -- stores the updated fp32 parameters and their tf32 shadow into EVERY rank's `p` / `p_tf32`.
// It replaces ncclAllReduce(5.74 MB) + the replicated Adam of the NCCL schedule (csrc/api.cu: run_step), whose all-reduce
// was fully exposed behind the backward pass (94 us of a 418 us step at 8 ranks, profiles/r1_dp8_diag.json).
//
// Cross-GPU protocol (no host involvement, no NCCL): every rank owns 2 x kMaxPeers arrival words.  A step has an epoch
// (local counter, identical on every rank because the ranks step in lockstep).
//   phase A  "my gradients are complete": true at kernel start (stream order); CTA 0 stores the epoch into word
//            [0][me] of every rank (st.release.sys).  Every CTA polls its OWN rank's words [0][*] (local L2) until all
//            ranks have arrived, then reads the peers' gradients.
//   phase B  "my stores into your p / p_tf32 are performed": each CTA fences (fence.sc.sys) after its last store and
//            counts itself done; the last CTA stores the epoch into word [1][me] of every rank and leaves.  The wait
//            for all ranks' [1][*] is a one-warp kernel at the head of the NEXT step's graph (peer_wait_kernel; any
//            other entry point of the library runs it first, api.cu: peer_quiesce): the next forward pass starts only
//            when every shard has landed here, no rank zeroes its `g` while a peer still reads it, and the wait
//            overlaps the next step's input staging instead of ending this kernel.
// Epochs only grow, so the words need no reset.  Every wait is bounded (~4 s) and traps: a dead peer surfaces as a
// launch failure of this rank, never as a hung GPU.  A CTA never waits for another CTA of its own grid (only for
// peers' kernels, which start independently of this rank), so no co-residency is assumed.
#include "common.cuh"
#include "kernels.h"

namespace vaeassoc {

namespace {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer gradients: system-scope loads (never served from a stale L1 line)
__device__ __forceinline__ float4 ld_sys_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys_f(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
// NVLS: the switch adds the W ranks' values of this address and returns one float4
__device__ __forceinline__ float4 multimem_ld_reduce_f4(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// NVLS: one store, replicated by the switch into every rank's copy
__device__ __forceinline__ void multimem_st_f4(float4* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// message passing needs release / acquire ordering only (data stores -> fence -> flag store): acq_rel, not sc
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void wait_epoch(const uint32_t* word, uint32_t epoch) {
  const long long t0 = clock64();
  while ((int32_t)(ld_relaxed_sys(word) - epoch) < 0) {
    __nanosleep(64);
    if (clock64() - t0 > 8000000000ll) __trap();    // ~4 s: a peer that never arrives must not hang this GPU
  }
}

template <int W, bool MC>
__global__ void __launch_bounds__(256) peer_adam_kernel(const PeerAdamArgs a) {
  __shared__ float s_lr_t;
  __shared__ uint32_t s_epoch, s_last;
  const AdamArgs& ad = a.adam;
  const bool stamp = a.tl != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  const unsigned long long t_entry = a.tl ? gtimer() : 0ull;
  const int me = a.rank;
  uint32_t* my_flags = a.flag_peer[me];
  if (threadIdx.x == 0) {
    s_epoch = a.sync[0] + 1u;                         // bumped by the last CTA of the previous step's kernel
    const double t = (double)(*ad.step_dev);          // already incremented by the finalize kernel: t >= 1
    const double b1t = pow((double)ad.beta1, t), b2t = pow((double)ad.beta2, t);
    s_lr_t = (float)((double)ad.lr * sqrt(1.0 - b2t) / (1.0 - b1t));
  }
  __syncthreads();
  const uint32_t epoch = s_epoch;
  // ---- phase A: announce, then wait for every rank's gradients ----
  if (blockIdx.x == 0 && threadIdx.x < W) st_release_sys(a.flag_peer[threadIdx.x] + me, epoch);
  if (threadIdx.x < W) {
    wait_epoch(my_flags + threadIdx.x, epoch);
    (void)ld_acquire_sys(my_flags + threadIdx.x);     // acquire: the peers' gradient stores happen-before our loads
  }
  __syncthreads();
  const unsigned long long t_a = stamp ? gtimer() : 0ull;

  const float lr_t = s_lr_t;
  const float b1 = ad.beta1, b2 = ad.beta2, ob1 = 1.0f - ad.beta1, ob2 = 1.0f - ad.beta2, eps = ad.eps;
  float4* __restrict__ m4 = reinterpret_cast<float4*>(ad.m);
  float4* __restrict__ v4 = reinterpret_cast<float4*>(ad.v);
  const float4* __restrict__ p4 = reinterpret_cast<const float4*>(ad.p);
  const bool shadow = ad.p_tf32 != nullptr;
  for (int64_t i = a.shard_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.shard_hi;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 g;
    if (MC) {
      g = multimem_ld_reduce_f4(reinterpret_cast<const float4*>(a.g_mc) + i);   // summed inside the NVSwitch
    } else {
      float4 gr[W];
#pragma unroll
      for (int r = 0; r < W; ++r) gr[r] = ld_sys_f4(reinterpret_cast<const float4*>(a.g_peer[r]) + i);   // W loads in flight
      g = gr[0];
#pragma unroll
      for (int r = 1; r < W; ++r) { g.x += gr[r].x; g.y += gr[r].y; g.z += gr[r].z; g.w += gr[r].w; }   // rank order: deterministic
    }
    float4 p = p4[i], m = m4[i], v = v4[i];
#define VAEASSOC_ADAM_LANE(c)                               \
    m.c = b1 * m.c + ob1 * g.c;                             \
    v.c = b2 * v.c + ob2 * g.c * g.c;                       \
    p.c = p.c - lr_t * m.c / (sqrtf(v.c) + eps);
    VAEASSOC_ADAM_LANE(x) VAEASSOC_ADAM_LANE(y) VAEASSOC_ADAM_LANE(z) VAEASSOC_ADAM_LANE(w)
#undef VAEASSOC_ADAM_LANE
    m4[i] = m; v4[i] = v;
    const float4 ps = make_float4(round_tf32(p.x), round_tf32(p.y), round_tf32(p.z), round_tf32(p.w));
    if (MC) {
      multimem_st_f4(reinterpret_cast<float4*>(a.p_mc) + i, p);
      if (shadow) multimem_st_f4(reinterpret_cast<float4*>(a.ptf_mc) + i, ps);
    } else {
#pragma unroll
      for (int r = 0; r < W; ++r) {
        reinterpret_cast<float4*>(a.p_peer[r])[i] = p;
        if (shadow) reinterpret_cast<float4*>(a.ptf_peer[r])[i] = ps;
      }
    }
  }
  // cost of the step = sum over ranks of the local cost slots (every rank computes the same sum in rank order)
  float cost = 0.0f;
  if (blockIdx.x == 0 && threadIdx.x == 0 && ad.cost_slot != nullptr) {
#pragma unroll
    for (int r = 0; r < W; ++r) cost += ld_sys_f(a.g_peer[r] + ad.n);
    if (ad.last_cost) *ad.last_cost = cost;
    if (ad.cost_hist) ad.cost_hist[((*ad.step_dev) - 1) % ad.hist_cap] = cost;
  }
  // ---- phase B: all stores of this rank performed -> tell every rank; leave only when every rank has told us ----
  __syncthreads();
  const unsigned long long t_d = stamp ? gtimer() : 0ull;
  if (threadIdx.x == 0) {
    fence_sys();                                       // cumulative over the CTA's stores (ordered by the barrier above)
    s_last = (atomicAdd(a.sync + 1, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  if (stamp) {
    const unsigned long long t_f = gtimer();
    atomicAdd(a.tl + 0, 1ull);
    atomicAdd(a.tl + 1, t_a - t_entry);                // waiting for the slowest rank's gradients
    atomicAdd(a.tl + 2, t_d - t_a);                    // reduce-scatter loads, Adam, all-gather stores (CTA 0)
    atomicAdd(a.tl + 3, t_f - t_d);                    // system fence of CTA 0
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) fence_sys();                 // the other CTAs' fenced stores -> before our release below
    __syncthreads();
    // (the wait for the OTHER ranks' phase-B words is the first node of the next step's graph -- peer_wait_kernel --
    // where it overlaps that step's input staging instead of ending this kernel)
    if (threadIdx.x < W) st_release_sys(a.flag_peer[threadIdx.x] + kMaxPeers + me, epoch);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (a.tl) atomicAdd(a.tl + 4, gtimer() - t_entry);   // whole kernel as seen by the last CTA
      // block 0 may not be the last CTA: the summed cost goes to the local slot here only if this CTA computed it; the
      // slot itself is rewritten by the rank's own finalize kernel every step, so publishing through last_cost /
      // cost_hist (above) is what the host reads
      a.sync[1] = 0u;
      a.sync[0] = epoch;
      __threadfence();
    }
  }
}

__global__ void peer_wait_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ sync, int world) {
  const uint32_t epoch = ld_relaxed_sys(sync);
  if ((int)threadIdx.x < world) {
    wait_epoch(flags + kMaxPeers + threadIdx.x, epoch);
    (void)ld_acquire_sys(flags + kMaxPeers + threadIdx.x);
  }
}

inline int peer_grid(int64_t n4) {
  int64_t b = (n4 + 255) / 256;
  int64_t waves = (b + kNumSMs - 1) / kNumSMs;
  if (waves > 4) waves = 4;
  if (waves < 1) waves = 1;
  return (int)(waves * kNumSMs);
}

}  // namespace

void launch_peer_wait(const uint32_t* flags, const uint32_t* sync, int world, cudaStream_t s) {
  peer_wait_kernel<<<1, 32, 0, s>>>(flags, sync, world);
}

void launch_peer_adam(const PeerAdamArgs& a, cudaStream_t s) {
  const int grid = peer_grid(a.shard_hi - a.shard_lo);
  const bool mc = a.g_mc != nullptr;
#define VAEASSOC_PEER_CASE(W)                                                        \
    case W:                                                                          \
      if (mc) peer_adam_kernel<W, true><<<grid, 256, 0, s>>>(a);                     \
      else peer_adam_kernel<W, false><<<grid, 256, 0, s>>>(a);                       \
      break;
  switch (a.world) {
    VAEASSOC_PEER_CASE(2) VAEASSOC_PEER_CASE(3) VAEASSOC_PEER_CASE(4) VAEASSOC_PEER_CASE(5)
    VAEASSOC_PEER_CASE(6) VAEASSOC_PEER_CASE(7) VAEASSOC_PEER_CASE(8)
    default: break;   // world 1 never reaches here (run_step)
  }
#undef VAEASSOC_PEER_CASE
}

}  // namespace vaeassoc
