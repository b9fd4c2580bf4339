// Device-resident synthetic paired-batch generator: replaces the reference's host input path
// (dataset.py:22-43 next_batch, utils.py:142-195 pickle -> arrays) which needs datasets we cannot download.
// Mirrors that path's value ranges and layout only; definition and oracle twin in oracle/synth.py.
//   image:  x[d] = 1{ sigmoid(2 c.P[:,d] - 3.5) > u1 } * (0.5 + 0.5 u2)      background exactly 0, lit in [0.5,1]
//   joint:  x[d] = (c.P[:,d] + 0.3 n) / sqrt(|P[:,d]|^2 + 0.09)              z-scored per column
// c ~ N(0, I_4) per global sample index, so a G-way shard reproduces the 1-GPU stream.
#include "common.cuh"
#include "kernels.h"
#include "philox.cuh"

namespace vaeassoc {

namespace {

constexpr int kCodeDim = 4;

__global__ void synth_projection_kernel(float* __restrict__ P, int modality, int n_input, uint32_t proj_seed) {
  const int n = kCodeDim * n_input;
  const int nblk = (n + 3) / 4;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += gridDim.x * blockDim.x) {
    float v[4];
    philox_normal4((uint64_t)b, 0u, 0u, proj_seed, TAG_PROJ + (uint32_t)modality, v);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (b * 4 + j < n) P[b * 4 + j] = v[j];
  }
}

__global__ void synth_invstd_kernel(const float* __restrict__ P, float* __restrict__ inv_std, int n_input) {
  for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < n_input; d += gridDim.x * blockDim.x) {
    float s = 0.09f;
#pragma unroll
    for (int k = 0; k < kCodeDim; ++k) s += P[k * n_input + d] * P[k * n_input + d];
    inv_std[d] = 1.0f / sqrtf(s);
  }
}

// one thread per Philox block: 2 pixels (binary) or 4 elements (Gaussian) of one row
template <bool BINARY>
__global__ void __launch_bounds__(256) synth_modality_kernel(float* __restrict__ x, int64_t ldx,
                                                             const float* __restrict__ P,
                                                             const float* __restrict__ inv_std, int modality,
                                                             int n_input, uint32_t data_seed, int64_t row0,
                                                             int64_t n_rows) {
  constexpr int PER = BINARY ? 2 : 4;
  const int nblk = (n_input + PER - 1) / PER;
  const int64_t total = n_rows * nblk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / nblk;
    const int b = (int)(i - r * nblk);
    const uint64_t row = (uint64_t)(row0 + r);
    float c[4];
    philox_normal4(row, 0u, 0u, data_seed, TAG_CODE, c);
    uint32_t w[4];
    philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), (uint32_t)b, (uint32_t)modality, data_seed,
                  BINARY ? (uint32_t)TAG_IMG : (uint32_t)TAG_JNT, w);
    float nrm[4];
    if (!BINARY) {
      philox_box_muller(w[0], w[1], nrm[0], nrm[1]);
      philox_box_muller(w[2], w[3], nrm[2], nrm[3]);
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int d = b * PER + j;
      if (d >= n_input) break;
      float proj = 0.f;
#pragma unroll
      for (int k = 0; k < kCodeDim; ++k) proj = fmaf(c[k], P[k * n_input + d], proj);
      float v;
      if (BINARY) {
        const float u1 = philox_u01(w[2 * j]), u2 = philox_u01(w[2 * j + 1]);
        const float p = 1.0f / (1.0f + expf(-(2.0f * proj - 3.5f)));
        v = (p > u1) ? 0.5f + 0.5f * u2 : 0.0f;
      } else {
        v = (proj + 0.3f * nrm[j]) * inv_std[d];
      }
      x[r * ldx + d] = v;
    }
  }
}

inline int grid_cap(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = 16 * kNumSMs;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

void launch_synth_projection(float* P, float* inv_std, int modality, int n_input, uint32_t proj_seed,
                             cudaStream_t s) {
  synth_projection_kernel<<<grid_cap(n_input), 256, 0, s>>>(P, modality, n_input, proj_seed);
  synth_invstd_kernel<<<grid_cap(n_input), 256, 0, s>>>(P, inv_std, n_input);
}

void launch_synth_modality(float* x, int64_t ldx, const float* P, const float* inv_std, int modality, int n_input,
                           int binary, uint32_t data_seed, int64_t row0, int64_t n_rows, cudaStream_t s) {
  if (binary) {
    const int64_t total = n_rows * ((n_input + 1) / 2);
    synth_modality_kernel<true><<<grid_cap(total), 256, 0, s>>>(x, ldx, P, inv_std, modality, n_input, data_seed,
                                                                row0, n_rows);
  } else {
    const int64_t total = n_rows * ((n_input + 3) / 4);
    synth_modality_kernel<false><<<grid_cap(total), 256, 0, s>>>(x, ldx, P, inv_std, modality, n_input, data_seed,
                                                                 row0, n_rows);
  }
}

}  // namespace vaeassoc
