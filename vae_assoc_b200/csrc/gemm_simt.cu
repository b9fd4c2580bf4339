// fp32 SIMT GEMMs (FFMA) for the dense layers: the fp32-exact precision mode (parity 1e-4) and the shapes the
// tcgen05 path does not serve (n_z-wide heads, K = n_z decoder input layer).
//   forward  y = act(x W + b)          vae_assoc.py:187-188,203-204,217-221,259-260,282-283,295-303
//   dgrad / wgrad                      autodiff of the above, vae_assoc.py:373-374
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
        if (gridDim.z > 1) atomicAdd(&g.C[(int64_t)gm * g.ldc + gn], v);
        else g.C[(int64_t)gm * g.ldc + gn] += v;
      }
    }
  }
  if (do_bsum) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < g.N) atomicAdd(&g.bias_grad[gn], bsum[j]);
    }
  }
}

// column sums of a [rows, cols] matrix (bias gradient = colsum of the upstream gradient) accumulated into out[cols].
// 32 x 8 threads: a warp reads 128 contiguous bytes of one row; 8 row-slabs per CTA reduce through shared memory;
// one fp32 RED per column per CTA.
constexpr int CS_ROWS = 256;
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int64_t ld, int rows, int cols,
                                                     float* __restrict__ out) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * CS_ROWS, r1 = min(rows, r0 + CS_ROWS);
  float acc = 0.f;
  if (c < cols)
    for (int r = r0 + ty; r < r1; r += 8) acc += X[(int64_t)r * ld + c];
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][tx];
    atomicAdd(out + c, t);
  }
}

inline dim3 grid_for(const GemmArgs& a, int splits) {
  return dim3((a.N + BN - 1) / BN, (a.M + BM - 1) / BM, splits);
}

}  // namespace

void launch_colsum(const float* X, int64_t ld, int rows, int cols, float* out, cudaStream_t s) {
  colsum_kernel<<<dim3((cols + 31) / 32, (rows + CS_ROWS - 1) / CS_ROWS), 256, 0, s>>>(X, ld, rows, cols, out);
}

void launch_gemm_nn_simt(const GemmArgs& a, cudaStream_t s) {
  gemm_simt_kernel<false, false, EPI_FWD><<<grid_for(a, 1), NT, 0, s>>>(a, a.K);
}

void launch_gemm_nt_simt(const GemmArgs& a, cudaStream_t s) {
  gemm_simt_kernel<false, true, EPI_DGRAD><<<grid_for(a, 1), NT, 0, s>>>(a, a.K);
}

void launch_gemm_tn_simt(const GemmArgs& a, cudaStream_t s) {
  // split the batch contraction so that the grid covers the 148 SMs a few times over
  const int tiles = ((a.N + BN - 1) / BN) * ((a.M + BM - 1) / BM);
  int splits = a.splitk > 0 ? a.splitk : 1;
  if (a.splitk <= 1) {
    const int want = (4 * kNumSMs + tiles - 1) / tiles;
    const int max_splits = (a.K + 4 * BK - 1) / (4 * BK);
    splits = max(1, min(want, max_splits));
  }
  int kps = (a.K + splits - 1) / splits;
  kps = ((kps + BK - 1) / BK) * BK;
  splits = (a.K + kps - 1) / kps;
  gemm_simt_kernel<true, false, EPI_WGRAD><<<grid_for(a, splits), NT, 0, s>>>(a, kps);
}

}  // namespace vaeassoc
