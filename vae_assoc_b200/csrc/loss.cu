// Fused reparameterisation + loss kernels (HBM-bound; 128-bit loads where rows allow, warp-shuffle reductions,
// deterministic two-pass cost reduction -- no floating-point atomics on the cost path).
//   z = mu + sqrt(exp(log sigma^2)) * eps                           vae_assoc.py:102-103
//   Bernoulli CE with the 1e-3 clamp                                vae_assoc.py:321-324
//   Gaussian reconstruction = tf.nn.l2_loss (batch SUM)             vae_assoc.py:327-328
//   prior KL                                                        vae_assoc.py:335-337
//   per-modality cost = mean_b(recon + KL) * weight                 vae_assoc.py:340
//   symmetric association KL, summed over the batch                 vae_assoc.py:346-366
//   total                                                           vae_assoc.py:368-371
// plus the analytic gradients tf.train.AdamOptimizer.minimize (:373-374) would derive.
#include "common.cuh"
#include "kernels.h"
#include "latent.cuh"
#include "philox.cuh"

namespace vaeassoc {

namespace {

// ---------------------------------------------------------------------------------------------------
// staging: copy caller rows into the library's padded input buffers (rounding to tf32 when they feed
// tcgen05 GEMMs) and produce this step's eps (copy of injected noise, or Philox).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stage_kernel(StageArgs a) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int m = 0; m < a.n_mod; ++m) {
    const float* __restrict__ src = a.src[m];
    if (src == nullptr) continue;
    float* __restrict__ dst = a.dst[m];
    const int ni = a.n_input[m];
    const int64_t sld = a.src_ld[m], dld = a.dst_ld[m];
    const bool vec = (ni % 4 == 0) && (sld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (vec) {
      const int q = ni / 4;
      const int64_t total = (int64_t)a.batch * q;
      for (int64_t i = tid; i < total; i += nthreads) {
        const int64_t r = i / q;
        const int c = (int)(i - r * q) * 4;
        const int64_t sr = a.row_index ? __ldg(a.row_index + r) : r;
        float4 v = __ldg(reinterpret_cast<const float4*>(src + sr * sld + c));
        if (a.round_tf32) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
        *reinterpret_cast<float4*>(dst + r * dld + c) = v;
      }
    } else {
      const int64_t total = (int64_t)a.batch * ni;
      for (int64_t i = tid; i < total; i += nthreads) {
        const int64_t r = i / ni;
        const int c = (int)(i - r * ni);
        const int64_t sr = a.row_index ? __ldg(a.row_index + r) : r;
        float v = __ldg(src + sr * sld + c);
        if (a.round_tf32) v = round_tf32(v);
        dst[r * dld + c] = v;
      }
    }
  }
  for (int z = 0; z < 8; ++z) {
    if (a.zero_ptr[z] == nullptr) continue;
    float4* __restrict__ zp = reinterpret_cast<float4*>(a.zero_ptr[z]);
    const int64_t n4 = a.zero_n[z] >> 2;
    for (int64_t i = tid; i < n4; i += nthreads) zp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (a.eps_dst != nullptr) {
    if (a.eps_src != nullptr) {
      const int64_t total = (int64_t)a.batch * a.n_z;
      for (int64_t i = tid; i < total; i += nthreads) a.eps_dst[i] = __ldg(a.eps_src + i);
    } else {
      const int nblk = (a.n_z + 3) / 4;
      const uint32_t step = (uint32_t)(*a.step_dev);
      const int64_t total = (int64_t)a.batch * nblk;
      for (int64_t i = tid; i < total; i += nthreads) {
        const int64_t r = i / nblk;
        const int b = (int)(i - r * nblk);
        float n[4];
        philox_normal4((uint64_t)(a.global_row0 + r), (uint32_t)b, step, a.eps_seed, TAG_EPS, n);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (b * 4 + j < a.n_z) a.eps_dst[r * a.n_z + b * 4 + j] = n[j];
      }
    }
  }
}

__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ dst, int64_t n_rows, int n_cols,
                                                            uint32_t seed, uint32_t tag, int64_t row0,
                                                            uint32_t step) {
  const int nblk = (n_cols + 3) / 4;
  const int64_t total = n_rows * nblk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nblk;
    const int b = (int)(i - r * nblk);
    float n[4];
    philox_normal4((uint64_t)(row0 + r), (uint32_t)b, step, seed, tag, n);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (b * 4 + j < n_cols) dst[r * n_cols + b * 4 + j] = n[j];
  }
}

// ---------------------------------------------------------------------------------------------------
// latent forward: one thread per row, looping over the n_z latent dimensions, so the per-row sums (prior KL,
// association KL over all modality pairs) are plain serial sums -- deterministic, no atomics.  The tensors are
// tiny ([B, 2 n_z] per modality); consecutive k of a row hit L1.
// ---------------------------------------------------------------------------------------------------
template <int NMOD>
__global__ void __launch_bounds__(256) latent_fwd_kernel(LatentArgs a) {
  __shared__ float red[32];
  float s_lat[NMOD];
  float s_assoc = 0.f;
#pragma unroll
  for (int m = 0; m < NMOD; ++m) s_lat[m] = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < a.batch; r += (int64_t)gridDim.x * blockDim.x) {
    float row_kl[NMOD], row_assoc;
    latent_fwd_row<NMOD, LoadPlain>(a, r, row_kl, row_assoc);
#pragma unroll
    for (int m = 0; m < NMOD; ++m) s_lat[m] += row_kl[m];
    s_assoc += row_assoc;
  }
  // deterministic block partials
#pragma unroll
  for (int m = 0; m < NMOD; ++m) {
    const float t = block_sum(s_lat[m], red);
    if (threadIdx.x == 0) a.partials[(int64_t)blockIdx.x * kCostSlots + 2 * m + 1] = t;
  }
  const float t = block_sum(s_assoc, red);
  if (threadIdx.x == 0) a.partials[(int64_t)blockIdx.x * kCostSlots + 8] = t;
}

template <int NMOD>
__global__ void __launch_bounds__(256) latent_bwd_kernel(LatentBwdArgs a) {
  const int nz = a.n_z;
  const int64_t total = (int64_t)a.batch * nz;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nz;
    const int k = (int)(i - r * nz);
#pragma unroll
    for (int m = 0; m < NMOD; ++m) {
      float dm, dl;
      latent_bwd_elem<LoadPlain>(a, m, r, k, dm, dl);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// reconstruction loss + d cost / d pre-activation.  One warp per row (rows are 784 / 147 floats), float4 loads
// when the row is 16-byte aligned, shuffle reduce per row, block partial per CTA.
// ---------------------------------------------------------------------------------------------------
// FAST (tf32 mode, tolerance 2e-3): ex2/lg2/rcp approximations -- the exact forms cost ~120 instructions per element
// and made this kernel issue-bound (37 us for 6.4 M elements) instead of HBM-bound
template <bool BINARY, bool FAST>
__device__ __forceinline__ float recon_elem(float x, float xh, float scale, float& da) {
  if (BINARY) {
    const float p = kCeEps + xh;                       // 1e-3 + x_hat
    const float q = (kCeEps + 1.0f) - xh;              // (1e-3 + 1) - x_hat, evaluation order of :323
    if (FAST) {
      da = scale * (__fdividef(1.0f - x, q) - __fdividef(x, p)) * xh * (1.0f - xh);
      return -(x * __logf(p) + (1.0f - x) * __logf(q));
    }
    da = scale * (-x / p + (1.0f - x) / q) * xh * (1.0f - xh);
    return -(x * logf(p) + (1.0f - x) * logf(q));
  } else {
    const float d = xh - x;
    da = scale * d;
    return 0.5f * d * d;
  }
}

constexpr int kReconVecIters = 8;   // rows up to 8 * 128 floats keep every load of the row in flight at once

template <bool BINARY, bool FAST>
__global__ void __launch_bounds__(256) recon_loss_kernel(ReconArgs a) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * warps_per_block;
  const int ni = a.n_input;
  const bool vec = (ni % 4 == 0) && (a.ldx % 4 == 0) && (a.ldxh % 4 == 0) && (a.ldda % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
  float block_acc = 0.f;
  for (int64_t r = warp0; r < a.batch; r += nwarps) {
    const float* __restrict__ x = a.x + r * a.ldx;
    const float* __restrict__ xh = a.xhat + r * a.ldxh;
    float* __restrict__ da = a.da ? a.da + r * a.ldda : nullptr;
    float acc = 0.f;
    if (vec && ni <= kReconVecIters * 128) {
      float4 xv[kReconVecIters], hv[kReconVecIters];
#pragma unroll
      for (int i = 0; i < kReconVecIters; ++i) {
        const int c = lane * 4 + i * 128;
        if (c < ni) {
          xv[i] = __ldg(reinterpret_cast<const float4*>(x + c));
          hv[i] = *reinterpret_cast<const float4*>(xh + c);
        }
      }
#pragma unroll
      for (int i = 0; i < kReconVecIters; ++i) {
        const int c = lane * 4 + i * 128;
        if (c < ni) {
          float4 d;
          acc += recon_elem<BINARY, FAST>(xv[i].x, hv[i].x, a.scale, d.x);
          acc += recon_elem<BINARY, FAST>(xv[i].y, hv[i].y, a.scale, d.y);
          acc += recon_elem<BINARY, FAST>(xv[i].z, hv[i].z, a.scale, d.z);
          acc += recon_elem<BINARY, FAST>(xv[i].w, hv[i].w, a.scale, d.w);
          if (da) {
            if (a.round_tf32) { d.x = round_tf32(d.x); d.y = round_tf32(d.y); d.z = round_tf32(d.z); d.w = round_tf32(d.w); }
            *reinterpret_cast<float4*>(da + c) = d;
          }
        }
      }
    } else if (vec) {
      for (int c = lane * 4; c < ni; c += 128) {
        const float4 xv = __ldg(reinterpret_cast<const float4*>(x + c));
        const float4 hv = *reinterpret_cast<const float4*>(xh + c);
        float4 d;
        acc += recon_elem<BINARY, FAST>(xv.x, hv.x, a.scale, d.x);
        acc += recon_elem<BINARY, FAST>(xv.y, hv.y, a.scale, d.y);
        acc += recon_elem<BINARY, FAST>(xv.z, hv.z, a.scale, d.z);
        acc += recon_elem<BINARY, FAST>(xv.w, hv.w, a.scale, d.w);
        if (da) {
          if (a.round_tf32) { d.x = round_tf32(d.x); d.y = round_tf32(d.y); d.z = round_tf32(d.z); d.w = round_tf32(d.w); }
          *reinterpret_cast<float4*>(da + c) = d;
        }
      }
    } else {
      for (int c = lane; c < ni; c += 32) {
        float d;
        acc += recon_elem<BINARY, FAST>(__ldg(x + c), xh[c], a.scale, d);
        if (da) da[c] = a.round_tf32 ? round_tf32(d) : d;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (BINARY && a.row_loss) a.row_loss[r] = acc;
      block_acc += acc;
    }
  }
  const float t = block_sum(block_acc, red);
  if (threadIdx.x == 0) a.partials[(int64_t)blockIdx.x * kCostSlots + a.slot] = t;
}

// ---------------------------------------------------------------------------------------------------
// finalize: fixed-order sum of the block partials -> scalars, local cost into the gradient buffer's spare slot
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) finalize_kernel(FinalizeArgs a) {
  __shared__ float sums[kCostSlots];
  // one warp per needed quantity (recon m, latent m, assoc): lanes stride the block partials, fixed-order shuffle sum
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int nq = 2 * a.n_mod + 1;
  for (int qi = warp; qi < nq; qi += nwarps) {
    const float* src; int nblk, slot, stride = kCostSlots, off;
    if (qi == 2 * a.n_mod) { src = a.partials_latent; nblk = a.blocks_latent; slot = 8; off = 8; }
    else if (qi & 1) { src = a.partials_latent; nblk = a.blocks_latent; slot = qi; off = qi; }
    else { src = a.partials_recon[qi >> 1]; nblk = a.blocks_recon[qi >> 1]; slot = qi; stride = a.stride_recon[qi >> 1]; off = a.off_recon[qi >> 1]; }
    float acc = 0.f;
    for (int b = lane; b < nblk; b += 32) acc += src[(int64_t)b * stride + off];
    acc = warp_sum(acc);
    if (lane == 0) sums[slot] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) finalize_combine(a, sums, a.advance);
}

inline int grid_for_elems(int64_t n, int per_block) {
  int64_t b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > kMaxPartialBlocks) b = kMaxPartialBlocks;
  return (int)b;
}

}  // namespace

void launch_stage(const StageArgs& a, cudaStream_t s) {
  int64_t total = (int64_t)a.batch * a.n_z;
  for (int m = 0; m < a.n_mod; ++m)
    if (a.src[m]) total += (int64_t)a.batch * a.n_input[m] / 4;
  const int grid = grid_for_elems(total, 256 * 2);
  stage_kernel<<<grid, 256, 0, s>>>(a);
}

void launch_philox_normal(float* dst, int64_t n_rows, int n_cols, uint32_t seed, uint32_t tag, int64_t row0,
                          uint32_t step, cudaStream_t s) {
  const int grid = grid_for_elems(n_rows * ((n_cols + 3) / 4), 256);
  philox_normal_kernel<<<grid, 256, 0, s>>>(dst, n_rows, n_cols, seed, tag, row0, step);
}

int launch_latent_fwd(const LatentArgs& a, cudaStream_t s) {
  const int grid = grid_for_elems((int64_t)a.batch, 256);
  switch (a.n_mod) {
    case 1: latent_fwd_kernel<1><<<grid, 256, 0, s>>>(a); break;
    case 2: latent_fwd_kernel<2><<<grid, 256, 0, s>>>(a); break;
    case 3: latent_fwd_kernel<3><<<grid, 256, 0, s>>>(a); break;
    default: latent_fwd_kernel<4><<<grid, 256, 0, s>>>(a); break;
  }
  return grid;
}

void launch_latent_bwd(const LatentBwdArgs& a, cudaStream_t s) {
  const int grid = grid_for_elems((int64_t)a.batch * a.n_z, 256);
  switch (a.n_mod) {
    case 1: latent_bwd_kernel<1><<<grid, 256, 0, s>>>(a); break;
    case 2: latent_bwd_kernel<2><<<grid, 256, 0, s>>>(a); break;
    case 3: latent_bwd_kernel<3><<<grid, 256, 0, s>>>(a); break;
    default: latent_bwd_kernel<4><<<grid, 256, 0, s>>>(a); break;
  }
}

int launch_recon_loss(const ReconArgs& a, cudaStream_t s) {
  // one warp per row, 8 warps per CTA; at most 8 x 148 CTAs (grid-stride beyond that)
  const int grid = grid_for_elems(a.batch, 8);
  if (a.binary && a.round_tf32) recon_loss_kernel<true, true><<<grid, 256, 0, s>>>(a);
  else if (a.binary) recon_loss_kernel<true, false><<<grid, 256, 0, s>>>(a);
  else recon_loss_kernel<false, false><<<grid, 256, 0, s>>>(a);
  return grid;
}

void launch_finalize(const FinalizeArgs& a, cudaStream_t s) { finalize_kernel<<<1, 256, 0, s>>>(a); }

}  // namespace vaeassoc
