// Multi-tensor Adam over ONE flat parameter buffer (HBM-bound: 28 B/param, +4 B for the tf32 shadow copy).
// TensorFlow's ApplyAdam formulation (what vae_assoc.py:373-374 runs; TensorFlow itself is not vendored):
//   lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
//   m <- beta1 m + (1-beta1) g ;  v <- beta2 v + (1-beta2) g^2 ;  p <- p - lr_t * m / (sqrt(v) + eps)
// (epsilon OUTSIDE the bias correction -- differs from torch.optim.Adam).  Padding gaps of the flat buffer have
// g = m = v = 0 and therefore stay exactly 0.  The reference launches 28 ApplyAdam kernels; this is one launch,
// 128-bit loads/stores, grid = a multiple of the 148 SMs.
#include "common.cuh"
#include "kernels.h"

namespace vaeassoc {

namespace {

__global__ void __launch_bounds__(256) adam_kernel(AdamArgs a) {
  __shared__ float s_lr_t;
  if (threadIdx.x == 0) {
    const double t = (double)(*a.step_dev);          // already incremented by the finalize kernel: t >= 1
    const double b1t = pow((double)a.beta1, t), b2t = pow((double)a.beta2, t);
    s_lr_t = (float)((double)a.lr * sqrt(1.0 - b2t) / (1.0 - b1t));
    if (blockIdx.x == 0 && a.cost_slot != nullptr) {
      const float c = *a.cost_slot;
      if (a.last_cost) *a.last_cost = c;
      if (a.cost_hist) a.cost_hist[((*a.step_dev) - 1) % a.hist_cap] = c;
    }
  }
  __syncthreads();
  const float lr_t = s_lr_t;
  const float b1 = a.beta1, b2 = a.beta2, ob1 = 1.0f - a.beta1, ob2 = 1.0f - a.beta2, eps = a.eps;
  const int64_t n4 = a.n >> 2;
  float4* __restrict__ p4 = reinterpret_cast<float4*>(a.p);
  float4* __restrict__ m4 = reinterpret_cast<float4*>(a.m);
  float4* __restrict__ v4 = reinterpret_cast<float4*>(a.v);
  const float4* __restrict__ g4 = reinterpret_cast<const float4*>(a.g);
  float4* __restrict__ s4 = reinterpret_cast<float4*>(a.p_tf32);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = p4[i], m = m4[i], v = v4[i];
    const float4 g = g4[i];
#define VAEASSOC_ADAM_LANE(c)                               \
    m.c = b1 * m.c + ob1 * g.c;                             \
    v.c = b2 * v.c + ob2 * g.c * g.c;                       \
    p.c = p.c - lr_t * m.c / (sqrtf(v.c) + eps);
    VAEASSOC_ADAM_LANE(x) VAEASSOC_ADAM_LANE(y) VAEASSOC_ADAM_LANE(z) VAEASSOC_ADAM_LANE(w)
#undef VAEASSOC_ADAM_LANE
    p4[i] = p; m4[i] = m; v4[i] = v;
    if (s4) s4[i] = make_float4(round_tf32(p.x), round_tf32(p.y), round_tf32(p.z), round_tf32(p.w));
    // the gradient buffer is an accumulator (TMA reduce-add, bias-gradient REDs): consumed here, it is handed back
    // cleared, so that the next step's graph needs no memset ahead of the tile kernel
    if (a.zero_g) reinterpret_cast<float4*>(a.zero_g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(256) round_copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                                         int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p = src[i];
    dst[i] = make_float4(round_tf32(p.x), round_tf32(p.y), round_tf32(p.z), round_tf32(p.w));
  }
}

__global__ void publish_cost_kernel(const float* cost_slot, float* last_cost) { *last_cost = *cost_slot; }

inline int adam_grid(int64_t n4) {
  int64_t b = (n4 + 255) / 256;
  // whole multiples of the SM count, up to 8 CTAs of 256 threads per SM
  int64_t waves = (b + kNumSMs - 1) / kNumSMs;
  if (waves > 8) waves = 8;
  if (waves < 1) waves = 1;
  return (int)(waves * kNumSMs);
}

}  // namespace

void launch_adam(const AdamArgs& a, cudaStream_t s) { adam_kernel<<<adam_grid(a.n >> 2), 256, 0, s>>>(a); }

void launch_round_copy(const float* src, float* dst, int64_t n, cudaStream_t s) {
  round_copy_kernel<<<adam_grid(n >> 2), 256, 0, s>>>(reinterpret_cast<const float4*>(src),
                                                      reinterpret_cast<float4*>(dst), n >> 2);
}

void launch_publish_cost(const float* cost_slot, float* last_cost, cudaStream_t s) {
  publish_cost_kernel<<<1, 1, 0, s>>>(cost_slot, last_cost);
}

}  // namespace vaeassoc
