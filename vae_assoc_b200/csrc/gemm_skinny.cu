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
// (thread per wide column, coalesced), the skinny operand's rows are broadcast from shared memory; CTA (x, y) reduces
// its slab of batch rows and stores the partial sums into the workspace ws[y][S][W] (+ column-sum partials); a second
// pass adds the slabs in ascending order: deterministic, no fp32 atomics.  bias_grad += colsum(B).
constexpr int ROWS = 32;
template <bool SKINNY_M>
__global__ void __launch_bounds__(128) skinny_wgrad_kernel(GemmArgs g, int rows_per_cta) {
  __shared__ float sS[ROWS][SK];
  const int S = SKINNY_M ? g.M : g.N;                       // skinny extent
  const int W = SKINNY_M ? g.N : g.M;                       // wide extent
  const float* skinny = SKINNY_M ? g.A : g.B;  const int64_t lds = SKINNY_M ? g.lda : g.ldb;
  const float* wide = SKINNY_M ? g.B : g.A;    const int64_t ldw = SKINNY_M ? g.ldb : g.lda;
  const int y0 = blockIdx.y * rows_per_cta, y1 = min(g.K, y0 + rows_per_cta);
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  float acc[SK];
#pragma unroll
  for (int s = 0; s < SK; ++s) acc[s] = 0.f;
  float colsum = 0.f, scs = 0.f;          // column sums of the wide operand (thread w) / of the skinny one (thread s)
  for (int r0 = y0; r0 < y1; r0 += ROWS) {
    const int r1 = min(y1, r0 + ROWS);
    __syncthreads();
    for (int i = threadIdx.x; i < (r1 - r0) * S; i += blockDim.x) {
      const int r = i / S, s = i - r * S;
      sS[r][s] = skinny[(int64_t)(r0 + r) * lds + s];
    }
    __syncthreads();
    if (w < W) {
#pragma unroll 8
      for (int r = r0; r < r1; ++r) {
        const float x = wide[(int64_t)r * ldw + w];
        colsum += x;
#pragma unroll
        for (int s = 0; s < SK; ++s)
          if (s < S) acc[s] = fmaf(sS[r - r0][s], x, acc[s]);
      }
    }
    if (!SKINNY_M && blockIdx.x == 0 && (int)threadIdx.x < S)
      for (int r = 0; r < r1 - r0; ++r) scs += sS[r][threadIdx.x];
  }
  float* __restrict__ part = g.ws + (int64_t)blockIdx.y * ((int64_t)S * W + max(S, W));
  if (w < W) {
#pragma unroll
    for (int s = 0; s < SK; ++s)
      if (s < S) part[(int64_t)s * W + w] = acc[s];
    if (SKINNY_M) part[(int64_t)S * W + w] = colsum;        // B is the wide operand when M is skinny
  }
  if (!SKINNY_M && blockIdx.x == 0 && (int)threadIdx.x < S) part[(int64_t)S * W + threadIdx.x] = scs;
}

// second pass: C[s, w] (or C[w, s]) += sum_y ws[y][s][w]; bias_grad += sum_y of the column-sum partials; y ascending
template <bool SKINNY_M>
__global__ void __launch_bounds__(256) skinny_wgrad_reduce_kernel(GemmArgs g, int ny) {
  const int S = SKINNY_M ? g.M : g.N, W = SKINNY_M ? g.N : g.M;
  const int64_t stride = (int64_t)S * W + max(S, W);
  const int nb = g.bias_grad ? g.N : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)S * W + nb; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    if (i < (int64_t)S * W) {
      for (int y = 0; y < ny; ++y) acc += g.ws[y * stride + i];
      const int s = (int)(i / W), w = (int)(i - (int64_t)s * W);
      float* dst = SKINNY_M ? g.C + (int64_t)s * g.ldc + w : g.C + (int64_t)w * g.ldc + s;
      *dst += acc;
    } else {
      const int n = (int)(i - (int64_t)S * W);
      for (int y = 0; y < ny; ++y) acc += g.ws[y * stride + (int64_t)S * W + n];
      g.bias_grad[n] += acc;
    }
  }
}

// batch rows per CTA of the skinny weight gradient: 32, or more so that at most 256 slabs are summed by the second pass
inline int skinny_wgrad_rows(int K) {
  int rpc = ROWS;
  if ((K + rpc - 1) / rpc > 256) rpc = (((K + 255) / 256) + ROWS - 1) / ROWS * ROWS;
  return rpc;
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
    const int W = skinny_m ? a.N : a.M, S = skinny_m ? a.M : a.N;
    const int rpc = skinny_wgrad_rows(a.K);
    const int ny = (a.K + rpc - 1) / rpc;
    dim3 grid((W + 127) / 128, ny);
    const int rgrid = (int)std::min<int64_t>(((int64_t)S * W + a.N + 255) / 256, 4 * kNumSMs);
    if (skinny_m) {
      skinny_wgrad_kernel<true><<<grid, 128, 0, s>>>(a, rpc);
      skinny_wgrad_reduce_kernel<true><<<rgrid, 256, 0, s>>>(a, ny);
    } else {
      skinny_wgrad_kernel<false><<<grid, 128, 0, s>>>(a, rpc);
      skinny_wgrad_reduce_kernel<false><<<rgrid, 256, 0, s>>>(a, ny);
    }
  }
}

// workspace floats of the skinny weight gradient (0 for the other kinds)
int64_t gemm_skinny_ws_floats(int kind, const GemmArgs& a) {
  if (kind != 2) return 0;
  const bool skinny_m = a.M <= SK;
  const int64_t W = skinny_m ? a.N : a.M, S = skinny_m ? a.M : a.N;
  const int rpc = skinny_wgrad_rows(a.K);
  return (int64_t)((a.K + rpc - 1) / rpc) * (S * W + std::max(S, W));
}

}  // namespace vaeassoc
