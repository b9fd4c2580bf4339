// Shared device helpers for libvaeassoc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vaeassoc {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// activation codes shared by the GEMM epilogues
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_SOFTPLUS = 2, ACT_SIGMOID = 3 };

__device__ __forceinline__ float round_tf32(float x) {
  // round-to-nearest (ties away) to a 10-bit mantissa: tcgen05 kind::tf32 ignores the low 13 bits, so operands
  // are rounded by their PRODUCER to keep the tensor-core path unbiased.
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ float softplus_f(float a) {
  // log(1 + e^a), stable for both tails (TF softplus: a for large a, e^a for very negative a)
  return fmaxf(a, 0.0f) + log1pf(expf(-fabsf(a)));
}

__device__ __forceinline__ float sigmoid_f(float a) { return 1.0f / (1.0f + expf(-a)); }

__device__ __forceinline__ float apply_act(int act, float v) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_SOFTPLUS: return softplus_f(v);
    case ACT_SIGMOID: return sigmoid_f(v);
    default: return v;
  }
}

// d act / d pre-activation, from the stored OUTPUT h: relu 1[h>0] (TF ReluGrad), softplus 1-exp(-h)
__device__ __forceinline__ float act_grad_from_output(int act, float h) {
  switch (act) {
    case ACT_RELU: return h > 0.0f ? 1.0f : 0.0f;
    case ACT_SOFTPLUS: return 1.0f - expf(-h);
    case ACT_SIGMOID: return h * (1.0f - h);
    default: return 1.0f;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum; every thread gets the result. `red` = shared float[32].
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0f;
  if (w == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

}  // namespace vaeassoc
