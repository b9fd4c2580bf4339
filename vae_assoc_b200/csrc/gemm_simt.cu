// fp32 SIMT GEMMs (FFMA) for the dense layers: the fp32-exact precision mode (parity 1e-4) and the shapes the
// tcgen05 path does not serve (n_z-wide heads, K = n_z decoder input layer).
//   forward  y = act(x W + b)          vae_assoc.py:187-188,203-204,217-221,259-260,282-283,295-303
//   dgrad / wgrad                      autodiff of the above, vae_assoc.py:373-374
// Every reduction here is DETERMINISTIC (no floating-point atomics): split-K weight gradients write their partial
// tiles into a workspace and a second pass sums them in a fixed order; column sums use per-CTA partials and a
// last-arriver pass in a fixed order.  Two runs of the fp32 path on the same inputs give bit-identical results
// (tests/test_gpu_parity.py::test_fp32_path_is_bit_reproducible) -- Adam's g / (|g| + 1e-8) turns one sign flip of a
// near-zero gradient into a different trajectory, so order noise in the fp32 path is not acceptable.
// Two tile shapes, 256 threads each, smem tiles stored k-major so the inner loop reads float4s per k:
//   128x128x8, 8x8 register micro-tile, 128-bit global loads where rows allow, register-prefetched and double-buffered
//              in shared memory (one barrier per k-tile): the layers of the fp32-exact mode with M, N >= 128;
//   64x64x16,  4x4 micro-tile, scalar loads: everything smaller.
// Both sum over k in ascending order with fmaf, so a contraction gives the same bits whichever shape serves it (split-K
// aside).  Any M/N/K/ld is accepted (bounds-checked loads; out-of-range elements contribute 0).
#include <cstdlib>

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

// ---- 128 x 128 x 8 form ------------------------------------------------------------------------------------------
constexpr int LBM = 128, LBN = 128, LBK = 8;

template <int EPI>
__device__ __forceinline__ void epi_element(const GemmArgs& g, int gm, int gn, float v) {
  if (EPI == EPI_FWD) {
    if (g.bias) v += g.bias[gn];
    if (g.aux) v *= act_grad_from_output(g.act, g.aux[(int64_t)gm * g.ldaux + gn]);
    else v = apply_act(g.act, v);
    if (g.round_out) v = round_tf32(v);
    g.C[(int64_t)gm * g.ldc + gn] = v;
  } else if (EPI == EPI_DGRAD) {
    if (g.aux) v *= act_grad_from_output(g.act, g.aux[(int64_t)gm * g.ldaux + gn]);
    if (g.round_out) v = round_tf32(v);
    g.C[(int64_t)gm * g.ldc + gn] = v;
  } else {
    if (gridDim.z > 1) g.ws[((int64_t)blockIdx.z * g.M + gm) * g.N + gn] = v;
    else g.C[(int64_t)gm * g.ldc + gn] += v;
  }
}

// one thread's four elements of a 128 x 8 operand tile.  CONTIG_MN: the operand is stored [K, X] (x contiguous): the
// thread takes k = tid / 32 and four consecutive x; else [X, K] (k contiguous): x = tid / 2 and four consecutive k.
template <bool CONTIG_MN>
__device__ __forceinline__ void load_tile4(const float* __restrict__ P, int64_t ld, bool vec_ok, int x0, int X, int k0, int kend,
                                           int tid, float (&r)[4]) {
  if (CONTIG_MN) {
    const int gk = k0 + (tid >> 5), gx = x0 + (tid & 31) * 4;
    const float* src = P + (int64_t)gk * ld + gx;
    if (gk < kend && gx + 3 < X && vec_ok) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src));
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = (gk < kend && gx + j < X) ? __ldg(src + j) : 0.f;
    }
  } else {
    const int gx = x0 + (tid >> 1), gk = k0 + (tid & 1) * 4;
    const float* src = P + (int64_t)gx * ld + gk;
    if (gx < X && gk + 3 < kend && vec_ok) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src));
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = (gx < X && gk + j < kend) ? __ldg(src + j) : 0.f;
    }
  }
}
template <bool CONTIG_MN>
__device__ __forceinline__ void store_tile4(float (*T)[LBM + 4], int tid, const float (&r)[4]) {
  if (CONTIG_MN) {
    *reinterpret_cast<float4*>(&T[tid >> 5][(tid & 31) * 4]) = make_float4(r[0], r[1], r[2], r[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) T[(tid & 1) * 4 + j][tid >> 1] = r[j];
  }
}

template <bool A_T, bool B_T, int EPI>
__global__ void __launch_bounds__(NT, 2) gemm_simt128_kernel(GemmArgs g, int k_per_split) {
  static_assert(LBM == LBN, "one tile-row type for both operands");
  __shared__ __align__(16) float As[2][LBK][LBM + 4];
  __shared__ __align__(16) float Bs[2][LBK][LBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * LBM, n0 = blockIdx.x * LBN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(g.K, kbeg + k_per_split);
  // 128-bit loads need 16-byte aligned rows (the k offsets of a split are multiples of LBK, tile offsets of 128)
  const bool vec_a = (g.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0);
  const bool vec_b = (g.ldb % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.B) & 15) == 0);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool do_bsum = (EPI == EPI_WGRAD) && g.bias_grad != nullptr && blockIdx.y == 0 && ty == 0;

  float ra[4], rb[4];
  // A: A_T = stored [K, M] (m contiguous); B: !B_T = stored [K, N] (n contiguous)
  load_tile4<A_T>(g.A, g.lda, vec_a, m0, g.M, kbeg, kend, tid, ra);
  load_tile4<!B_T>(g.B, g.ldb, vec_b, n0, g.N, kbeg, kend, tid, rb);
  store_tile4<A_T>(As[0], tid, ra);
  store_tile4<!B_T>(Bs[0], tid, rb);
  __syncthreads();
  int buf = 0;
  for (int k0 = kbeg; k0 < kend; k0 += LBK, buf ^= 1) {
    const bool more = k0 + LBK < kend;
    if (more) {                        // next k-tile: global -> registers while this one is consumed from shared memory
      load_tile4<A_T>(g.A, g.lda, vec_a, m0, g.M, k0 + LBK, kend, tid, ra);
      load_tile4<!B_T>(g.B, g.ldb, vec_b, n0, g.N, k0 + LBK, kend, tid, rb);
    }
#pragma unroll
    for (int k = 0; k < LBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (do_bsum) {
#pragma unroll
        for (int j = 0; j < 8; ++j) bsum[j] += bv[j];
      }
    }
    if (more) {
      store_tile4<A_T>(As[buf ^ 1], tid, ra);
      store_tile4<!B_T>(Bs[buf ^ 1], tid, rb);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= g.N) continue;
      epi_element<EPI>(g, gm, gn, acc[i][j]);
    }
  }
  if (do_bsum) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
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

inline bool use_big_tiles(const GemmArgs& a) {
  static const bool off = getenv("VAEASSOC_SIMT_SMALL_TILES") != nullptr;
  return !off && a.M >= LBM && a.N >= LBN;
}
inline dim3 grid_for(const GemmArgs& a, int splits) {
  if (use_big_tiles(a)) return dim3((a.N + LBN - 1) / LBN, (a.M + LBM - 1) / LBM, splits);
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
  if (use_big_tiles(a)) gemm_simt128_kernel<false, false, EPI_FWD><<<grid_for(a, 1), NT, 0, s>>>(a, a.K);
  else gemm_simt_kernel<false, false, EPI_FWD><<<grid_for(a, 1), NT, 0, s>>>(a, a.K);
}

void launch_gemm_nt_simt(const GemmArgs& a, cudaStream_t s) {
  if (use_big_tiles(a)) gemm_simt128_kernel<false, true, EPI_DGRAD><<<grid_for(a, 1), NT, 0, s>>>(a, a.K);
  else gemm_simt_kernel<false, true, EPI_DGRAD><<<grid_for(a, 1), NT, 0, s>>>(a, a.K);
}

namespace {
// split the batch contraction so that the grid covers the 148 SMs a few times over
void tn_splits(const GemmArgs& a, int* splits_out, int* kps_out) {
  const bool big = use_big_tiles(a);
  const int bm = big ? LBM : BM, bn = big ? LBN : BN, bk = big ? 2 * LBK : BK;   // (k ranges in multiples of 16 either way)
  const int tiles = ((a.N + bn - 1) / bn) * ((a.M + bm - 1) / bm);
  int splits = a.splitk > 0 ? a.splitk : 1;
  if (a.splitk <= 1) {
    const int want = ((big ? 2 : 4) * kNumSMs + tiles - 1) / tiles;
    const int max_splits = (a.K + 4 * bk - 1) / (4 * bk);
    splits = max(1, min(want, max_splits));
  }
  int kps = (a.K + splits - 1) / splits;
  kps = ((kps + bk - 1) / bk) * bk;
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
  if (use_big_tiles(a)) gemm_simt128_kernel<true, false, EPI_WGRAD><<<grid_for(a, splits), NT, 0, s>>>(a, kps);
  else gemm_simt_kernel<true, false, EPI_WGRAD><<<grid_for(a, splits), NT, 0, s>>>(a, kps);
  if (splits > 1) {
    const int64_t total = (int64_t)a.M * a.N + (a.bias_grad ? a.N : 0);
    const int64_t want = (total + 255) / 256;
    const int grid = (int)(want < 8 * kNumSMs ? want : 8 * kNumSMs);
    splitk_reduce_kernel<<<grid, 256, 0, s>>>(a.ws, splits, a.M, a.N, a.C, a.ldc, a.bias_grad);
  }
}

}  // namespace vaeassoc
