// fp32 SIMT GEMMs (FFMA) for the dense layers: the fp32-exact precision mode (parity 1e-4) and the shapes the
// tcgen05 path does not serve (n_z-wide heads, K = n_z decoder input layer).
//   forward  y = act(x W + b)          vae_assoc.py:187-188,203-204,217-221,259-260,282-283,295-303
//   dgrad / wgrad                      autodiff of the above, vae_assoc.py:373-374
// Every reduction here is DETERMINISTIC (no floating-point atomics): split-K weight gradients write their partial
// tiles into a workspace and a second pass sums them in a fixed order; column sums use per-CTA partials and a
// last-arriver pass in a fixed order.  Two runs of the fp32 path on the same inputs give bit-identical results
// (tests/test_gpu_parity.py::test_fp32_path_is_bit_reproducible) -- Adam's g / (|g| + 1e-8) turns one sign flip of a
// near-zero gradient into a different trajectory, so order noise in the fp32 path is not acceptable.
// 64x64x16 CTA tile, 256 threads, 4x4 register micro-tile, smem tiles stored k-major so the inner loop reads
// two float4 per k.  Any M/N/K/ld is accepted (bounds-checked loads; out-of-range elements contribute 0).
#include "common.cuh"
#include "kernels.h"

namespace vaeassoc {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

enum Epi : int { EPI_FWD = 0, EPI_DGRAD = 1, EPI_WGRAD = 2 };

// A_T: A stored [K, M] (m contiguous)   else [M, K] (k contiguous)
// B_T: B stored [N, K] (k contiguous)   else [K, N] (n contiguous)
template <bool A_T, bool B_T, int EPI>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(GemmArgs g, int k_per_split) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(g.K, kbeg + k_per_split);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};   // wgrad: column sums of B (bias gradient), rows ty==0 of m-tile 0
  const bool do_bsum = (EPI == EPI_WGRAD) && g.bias_grad != nullptr && blockIdx.y == 0 && ty == 0;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- global -> shared ------------------------------------------------------------------------
    if (A_T) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = (tid >> 6) + 4 * i, m = tid & 63;
        const int gk = k0 + k, gm = m0 + m;
        As[k][m] = (gk < kend && gm < g.M) ? g.A[(int64_t)gk * g.lda + gm] : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int m = (tid >> 4) + 16 * i, k = tid & 15;
        const int gk = k0 + k, gm = m0 + m;
        As[k][m] = (gk < kend && gm < g.M) ? g.A[(int64_t)gm * g.lda + gk] : 0.f;
      }
    }
    if (B_T) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = (tid >> 4) + 16 * i, k = tid & 15;
        const int gk = k0 + k, gn = n0 + n;
        Bs[k][n] = (gk < kend && gn < g.N) ? g.B[(int64_t)gn * g.ldb + gk] : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = (tid >> 6) + 4 * i, n = tid & 63;
        const int gk = k0 + k, gn = n0 + n;
        Bs[k][n] = (gk < kend && gn < g.N) ? g.B[(int64_t)gk * g.ldb + gn] : 0.f;
      }
    }
    __syncthreads();
    // ---- FFMA -------------------------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (do_bsum) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bsum[j] += bv[j];
      }
    }
    __syncthreads();
  }

  // ---- epilogue -------------------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      float v = acc[i][j];
      if (EPI == EPI_FWD) {
        if (g.bias) v += g.bias[gn];
        // with `aux` the forward form computes a gradient: v * act'(aux) (deconv backward, conv.cu) instead of act(v)
        if (g.aux) v *= act_grad_from_output(g.act, g.aux[(int64_t)gm * g.ldaux + gn]);
        else v = apply_act(g.act, v);
        if (g.round_out) v = round_tf32(v);
        g.C[(int64_t)gm * g.ldc + gn] = v;
      } else if (EPI == EPI_DGRAD) {
        if (g.aux) v *= act_grad_from_output(g.act, g.aux[(int64_t)gm * g.ldaux + gn]);
        if (g.round_out) v = round_tf32(v);
        g.C[(int64_t)gm * g.ldc + gn] = v;
      } else {
        // split-K: partial tile of split z into the workspace [z][M][N]; splitk_reduce_kernel adds them up in order
        if (gridDim.z > 1) g.ws[((int64_t)blockIdx.z * g.M + gm) * g.N + gn] = v;
        else g.C[(int64_t)gm * g.ldc + gn] += v;
      }
    }
  }
  if (do_bsum) {
    // exactly one thread of the launch owns (split z, column gn): a plain store / add, no atomics
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      if (gridDim.z > 1) g.ws[(int64_t)gridDim.z * g.M * g.N + (int64_t)blockIdx.z * g.N + gn] = bsum[j];
      else g.bias_grad[gn] += bsum[j];
    }
  }
}

// second pass of a split contraction: C[m, n] += sum_z ws[z][m][n] and bias_grad[n] += sum_z wsb[z][n], z ascending
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int splits, int M, int N,
                                                            float* __restrict__ C, int64_t ldc,
                                                            float* __restrict__ bias_grad) {
  const int64_t total = (int64_t)M * N;
  const float* __restrict__ wsb = ws + (int64_t)splits * total;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total + (bias_grad ? N : 0);
       i += (int64_t)gridDim.x * blockDim.x) {
    if (i < total) {
      float acc = 0.f;
      for (int z = 0; z < splits; ++z) acc += ws[(int64_t)z * total + i];
      const int64_t m = i / N;
      C[m * ldc + (i - m * N)] += acc;
    } else {
      const int n = (int)(i - total);
      float acc = 0.f;
      for (int z = 0; z < splits; ++z) acc += wsb[(int64_t)z * N + n];
      bias_grad[n] += acc;
    }
  }
}

// column sums of a [rows, cols] matrix (bias gradient = colsum of the upstream gradient) accumulated into out[cols].
// 32 x 8 threads: a warp reads 128 contiguous bytes of one row; 8 row-slabs per CTA reduce through shared memory.
// Deterministic: CTA (x, y) stores its partial into ws[y][32 x + lane]; the LAST CTA of column block x to arrive
// (ticket counter behind the partials) adds the gridDim.y partials in ascending y and updates out -- no fp32 atomics.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int64_t ld, int rows, int cols,
                                                     int rows_per_cta, float* __restrict__ out, float* __restrict__ ws) {
  __shared__ float part[8][33];
  __shared__ unsigned s_last;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  const int cols_pad = gridDim.x * 32;
  float acc = 0.f;
  if (c < cols)
    for (int r = r0 + ty; r < r1; r += 8) acc += X[(int64_t)r * ld + c];
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][tx];
    if (gridDim.y == 1) {
      if (c < cols) out[c] += t;
      return;
    }
    ws[(int64_t)blockIdx.y * cols_pad + c] = t;
    __threadfence();
  }
  if (gridDim.y == 1) return;
  __syncthreads();
  unsigned* tickets = reinterpret_cast<unsigned*>(ws + (int64_t)gridDim.y * cols_pad);
  if (threadIdx.x == 0) s_last = (atomicAdd(tickets + blockIdx.x, 1u) == gridDim.y - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // fixed order: slab ty sums partials ty, ty + 8, ... ascending; then the 8 slabs ascending
  float t = 0.f;
  for (int y = ty; y < (int)gridDim.y; y += 8) t += __ldcg(ws + (int64_t)y * cols_pad + c);
  part[ty][tx] = t;
  __syncthreads();
  if (ty == 0) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += part[i][tx];
    if (c < cols) out[c] += tot;
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;     // self-cleaning: the next launch (next step / graph replay) starts at 0
}

// rows per CTA: 256, or more when that would give more than 256 partials per column
inline int colsum_rows_per_cta(int64_t rows) {
  int64_t rpc = 256;
  if ((rows + rpc - 1) / rpc > 256) rpc = ((rows + 255) / 256 + 7) / 8 * 8;
  return (int)rpc;
}

inline dim3 grid_for(const GemmArgs& a, int splits) {
  return dim3((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, splits);
}

}  // namespace

int64_t colsum_ws_floats(int64_t rows, int cols) {
  const int rpc = colsum_rows_per_cta(rows);
  const int64_t ny = (rows + rpc - 1) / rpc, nx = (cols + 31) / 32;
  return ny * nx * 32 + nx;        // partials [ny][32 nx] + one ticket per column block (zero-initialised by the owner)
}

void launch_colsum(const float* X, int64_t ld, int64_t rows, int cols, float* out, float* ws, cudaStream_t s) {
  const int rpc = colsum_rows_per_cta(rows);
  colsum_kernel<<<dim3((cols + 31) / 32, (unsigned)((rows + rpc - 1) / rpc)), 256, 0, s>>>(X, ld, (int)rows, cols, rpc, out, ws);
}

void launch_gemm_nn_simt(const GemmArgs& a, cudaStream_t s) {
  gemm_simt_kernel<false, false, EPI_FWD><<<grid_for(a, 1), NT, 0, s>>>(a, a.K);
}

void launch_gemm_nt_simt(const GemmArgs& a, cudaStream_t s) {
  gemm_simt_kernel<false, true, EPI_DGRAD><<<grid_for(a, 1), NT, 0, s>>>(a, a.K);
}

namespace {
// split the batch contraction so that the grid covers the 148 SMs a few times over
void tn_splits(const GemmArgs& a, int* splits_out, int* kps_out) {
  const int tiles = ((a.N + BN - 1) / BN) * ((a.M + BM - 1) / BM);
  int splits = a.splitk > 0 ? a.splitk : 1;
  if (a.splitk <= 1) {
    const int want = (4 * kNumSMs + tiles - 1) / tiles;
    const int max_splits = (a.K + 4 * BK - 1) / (4 * BK);
    splits = max(1, min(want, max_splits));
  }
  int kps = (a.K + splits - 1) / splits;
  kps = ((kps + BK - 1) / BK) * BK;
  *splits_out = (a.K + kps - 1) / kps;
  *kps_out = kps;
}
}  // namespace

// workspace of the split weight gradient: [splits][M][N] partial tiles + [splits][N] partial bias sums
int64_t gemm_tn_simt_ws_floats(const GemmArgs& a) {
  int splits, kps;
  tn_splits(a, &splits, &kps);
  return splits > 1 ? (int64_t)splits * ((int64_t)a.M * a.N + a.N) : 0;
}

void launch_gemm_tn_simt(const GemmArgs& a, cudaStream_t s) {
  int splits, kps;
  tn_splits(a, &splits, &kps);
  if (splits > 1 && a.ws == nullptr) { splits = 1; kps = ((a.K + BK - 1) / BK) * BK; }   // no workspace: one ordered pass
  gemm_simt_kernel<true, false, EPI_WGRAD><<<grid_for(a, splits), NT, 0, s>>>(a, kps);
  if (splits > 1) {
    const int64_t total = (int64_t)a.M * a.N + (a.bias_grad ? a.N : 0);
    const int64_t want = (total + 255) / 256;
    const int grid = (int)(want < 8 * kNumSMs ? want : 8 * kNumSMs);
    splitk_reduce_kernel<<<grid, 256, 0, s>>>(a.ws, splits, a.M, a.N, a.C, a.ldc, a.bias_grad);
  }
}

}  // namespace vaeassoc
