// HBM-bound "skinny" dense contractions: the n_z-wide layers of the assoc-VAE (n_z = 4 in the reference config):
//   heads      [B, r2] . [r2, 2 n_z]          vae_assoc.py:217-221   (forward N = 8, dgrad K = 8, wgrad N = 8)
//   decoder-1  [B, n_z] . [n_z, r1]            vae_assoc.py:257-260   (forward K = 4, dgrad N = 4, wgrad M = 4)
// A 64x64 GEMM tile wastes 8-16x of its lanes on these; each kernel below instead streams the one large operand
// exactly once with coalesced loads and keeps the skinny operand in registers / shared memory.  Algorithmic bytes
// = 4 * (large operand + output); no tensor cores (intensity 2-4 FLOP/B).
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace vaeassoc {

namespace {

constexpr int SK = 16;   // largest "skinny" extent served

__device__ __forceinline__ float epilogue(const GemmArgs& g, float v, int m, int n) {
  if (g.bias) v += __ldg(g.bias + n);
  v = apply_act(g.aux ? ACT_NONE : g.act, v);
  if (g.aux) v *= act_grad_from_output(g.act, g.aux[(int64_t)m * g.ldaux + n]);
  if (g.round_out) v = round_tf32(v);
  return v;
}

// ---- (a)/(c): C[M, N<=16] = A[M,K] . B ;  B is [K,N] (B_T = false) or [N,K] (B_T = true).  One warp per row:
// lanes stride K (coalesced reads of the A row), N partial sums per lane, shuffle reduce.
template <bool B_T>
__global__ void __launch_bounds__(256) skinny_out_kernel(GemmArgs g) {
  extern __shared__ float sB[];                    // [K][NP]: row stride NP = N|1 (odd) keeps the lane-strided reads conflict-free
  const int N = g.N, K = g.K, NP = g.N | 1;
  for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
    const int k = i / N, n = i - k * N;
    sB[k * NP + n] = B_T ? g.B[(int64_t)n * g.ldb + k] : g.B[(int64_t)k * g.ldb + n];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int m = blockIdx.x * wpb + warp; m < g.M; m += gridDim.x * wpb) {
    float acc[SK];
#pragma unroll
    for (int n = 0; n < SK; ++n) acc[n] = 0.f;
    const float* __restrict__ a = g.A + (int64_t)m * g.lda;
    int k = lane;
    for (; k + 96 < K; k += 128) {          // four independent 128-byte row segments in flight per warp
      const float a0 = a[k], a1 = a[k + 32], a2 = a[k + 64], a3 = a[k + 96];
#pragma unroll
      for (int n = 0; n < SK; ++n)
        if (n < N)
          acc[n] = fmaf(a3, sB[(k + 96) * NP + n],
                        fmaf(a2, sB[(k + 64) * NP + n], fmaf(a1, sB[(k + 32) * NP + n], fmaf(a0, sB[k * NP + n], acc[n]))));
    }
    for (; k < K; k += 32) {
      const float av = a[k];
#pragma unroll
      for (int n = 0; n < SK; ++n)
        if (n < N) acc[n] = fmaf(av, sB[k * NP + n], acc[n]);
    }
#pragma unroll
    for (int n = 0; n < SK; ++n)
      if (n < N) acc[n] = warp_sum(acc[n]);
    if (lane == 0) {
#pragma unroll
      for (int n = 0; n < SK; ++n)
        if (n < N) g.C[(int64_t)m * g.ldc + n] = epilogue(g, acc[n], m, n);
    }
  }
}

// ---- (b)/(d): C[M,N] = A[M, K<=16] . B.  Thread <-> output column n with B[:, n] held in registers; the CTA walks
// a chunk of rows, reading each A row (K floats, warp-uniform -> one broadcast transaction) and writing C coalesced.
constexpr int IN_ROWS = 16;
template <bool B_T>
__global__ void __launch_bounds__(128) skinny_in_kernel(GemmArgs g) {
  const int K = g.K;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m0 = blockIdx.y * IN_ROWS, m1 = min(g.M, m0 + IN_ROWS);
  if (n >= g.N) return;
  float b[SK];
#pragma unroll
  for (int k = 0; k < SK; ++k)
    b[k] = (k < K) ? (B_T ? __ldg(g.B + (int64_t)n * g.ldb + k) : __ldg(g.B + (int64_t)k * g.ldb + n)) : 0.f;
#pragma unroll 4
  for (int m = m0; m < m1; ++m) {
    const float* __restrict__ a = g.A + (int64_t)m * g.lda;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < SK; ++k)
      if (k < K) acc = fmaf(__ldg(a + k), b[k], acc);
    g.C[(int64_t)m * g.ldc + n] = epilogue(g, acc, m, n);
  }
}

// ---- (e)/(f): wgrad C[M,N] += A[K,M]^T . B[K,N] with min(M,N) <= 16, K = batch.  The wide operand is streamed once
// (thread per wide column, coalesced), the skinny operand's rows are broadcast from shared memory; each CTA reduces
// ROWS batch rows and issues one fp32 RED per output element.  bias_grad += colsum(B).
constexpr int ROWS = 32;
template <bool SKINNY_M>
__global__ void __launch_bounds__(128) skinny_wgrad_kernel(GemmArgs g) {
  __shared__ float sS[ROWS][SK];
  const int S = SKINNY_M ? g.M : g.N;                       // skinny extent
  const int W = SKINNY_M ? g.N : g.M;                       // wide extent
  const float* skinny = SKINNY_M ? g.A : g.B;  const int64_t lds = SKINNY_M ? g.lda : g.ldb;
  const float* wide = SKINNY_M ? g.B : g.A;    const int64_t ldw = SKINNY_M ? g.ldb : g.lda;
  const int r0 = blockIdx.y * ROWS, r1 = min(g.K, r0 + ROWS);
  for (int i = threadIdx.x; i < (r1 - r0) * S; i += blockDim.x) {
    const int r = i / S, s = i - r * S;
    sS[r][s] = skinny[(int64_t)(r0 + r) * lds + s];
  }
  __syncthreads();
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[SK];
#pragma unroll
  for (int s = 0; s < SK; ++s) acc[s] = 0.f;
  float colsum = 0.f;
  if (w < W) {
#pragma unroll 8
    for (int r = r0; r < r1; ++r) {
      const float x = wide[(int64_t)r * ldw + w];
      colsum += x;
#pragma unroll
      for (int s = 0; s < SK; ++s)
        if (s < S) acc[s] = fmaf(sS[r - r0][s], x, acc[s]);
    }
#pragma unroll
    for (int s = 0; s < SK; ++s) {
      if (s < S) {
        float* dst = SKINNY_M ? g.C + (int64_t)s * g.ldc + w : g.C + (int64_t)w * g.ldc + s;
        atomicAdd(dst, acc[s]);
      }
    }
    // bias gradient = column sums of B: B is the wide operand when M is skinny
    if (SKINNY_M && g.bias_grad) atomicAdd(g.bias_grad + w, colsum);
  }
  if (!SKINNY_M && g.bias_grad && blockIdx.x == 0) {
    // B is the skinny operand: its column sums over this CTA's rows, from shared memory
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      float t = 0.f;
      for (int r = 0; r < r1 - r0; ++r) t += sS[r][s];
      atomicAdd(g.bias_grad + s, t);
    }
  }
}

inline int cap_grid(int64_t blocks) {
  const int64_t cap = 16 * kNumSMs;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

bool skinny_supported(int kind, const GemmArgs& a) {
  switch (kind) {
    case 0:
    case 1: return (a.N <= SK && (int64_t)a.K * (a.N | 1) * 4 <= 48 * 1024) || a.K <= SK;
    default: return a.M <= SK || a.N <= SK;
  }
}

void launch_gemm_skinny(int kind, const GemmArgs& a, cudaStream_t s) {
  if (kind == 0 || kind == 1) {
    if (a.K <= SK) {
      dim3 grid((a.N + 127) / 128, (a.M + IN_ROWS - 1) / IN_ROWS);
      if (kind == 0) skinny_in_kernel<false><<<grid, 128, 0, s>>>(a);
      else skinny_in_kernel<true><<<grid, 128, 0, s>>>(a);
    } else {
      // each CTA stages B in shared memory once and then walks rows: 4 CTAs per SM, grid-stride over the rows
      const int grid = (int)std::min<int64_t>((a.M + 7) / 8, 4 * kNumSMs);
      const size_t smem = (size_t)a.K * (a.N | 1) * 4;
      if (kind == 0) skinny_out_kernel<false><<<grid, 256, smem, s>>>(a);
      else skinny_out_kernel<true><<<grid, 256, smem, s>>>(a);
    }
  } else {
    const bool skinny_m = a.M <= SK;
    const int W = skinny_m ? a.N : a.M;
    dim3 grid((W + 127) / 128, (a.K + ROWS - 1) / ROWS);
    if (skinny_m) skinny_wgrad_kernel<true><<<grid, 128, 0, s>>>(a);
    else skinny_wgrad_kernel<false><<<grid, 128, 0, s>>>(a);
  }
}

}  // namespace vaeassoc
