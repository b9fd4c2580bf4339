// Row-level arithmetic of the latent stage (reparameterisation, prior KL, symmetric association KL and their
// gradients; vae_assoc.py:102-103, 335-337, 346-366), shared by the stand-alone kernels (loss.cu) and by the
// elementwise tasks of the persistent tile kernel (gemm_group.cu), so that both paths produce the same bits.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace vaeassoc {

// LOAD: how global memory is read -- plain loads in a stand-alone kernel (its inputs were written by earlier launches),
// L1-bypassing loads inside the persistent kernel (the inputs were written by other SMs of the SAME launch)
struct LoadPlain { __device__ __forceinline__ static float ld(const float* p) { return *p; } };
struct LoadCg { __device__ __forceinline__ static float ld(const float* p) { return __ldcg(p); } };

constexpr int kLatentRegs = 4;    // latent dimensions loaded per batch (every load of the batch in flight at once)

// one batch row of the latent forward; accumulates the row's prior-KL sums and association-KL sum.  All inputs of the
// row are loaded BEFORE the first store: the stores to z / gstat may alias the inputs as far as the compiler knows, so a
// load placed after them waits for them -- with one L2 round trip (~1 us inside the persistent kernel) per latent
// dimension the row cost 4-8 us instead of ~1.5
template <int NMOD, typename LOAD>
__device__ __forceinline__ void latent_fwd_row(const LatentArgs& a, int64_t r, float (&row_kl)[NMOD], float& row_assoc,
                                               bool add_head_bias = false) {
  const int nz = a.n_z;
#pragma unroll
  for (int m = 0; m < NMOD; ++m) row_kl[m] = 0.f;
  row_assoc = 0.f;
  for (int k0 = 0; k0 < nz; k0 += kLatentRegs) {
    const int kn = min(kLatentRegs, nz - k0);
    float e_[kLatentRegs], mu_[NMOD][kLatentRegs], lv_[NMOD][kLatentRegs];
#pragma unroll
    for (int i = 0; i < kLatentRegs; ++i) {
      if (i < kn) {
        e_[i] = LOAD::ld(a.eps + r * nz + k0 + i);
#pragma unroll
        for (int m = 0; m < NMOD; ++m) {
          mu_[m][i] = LOAD::ld(a.heads[m] + r * 2 * nz + k0 + i);
          lv_[m][i] = LOAD::ld(a.heads[m] + r * 2 * nz + nz + k0 + i);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kLatentRegs; ++i) {
      if (i >= kn) continue;
      const int k = k0 + i;
      const float e = e_[i];
      float mu[NMOD], lv[NMOD], ex[NMOD];
#pragma unroll
      for (int m = 0; m < NMOD; ++m) {
        mu[m] = mu_[m][i];
        lv[m] = lv_[m][i];
        if (add_head_bias) {                             // split-K heads: sums of the k ranges + the layer's bias; the
          mu[m] += __ldg(a.head_bias[m] + k);            // completed row goes back (probes, latent backward)
          lv[m] += __ldg(a.head_bias[m] + nz + k);
          float* hw = const_cast<float*>(a.heads[m]);
          hw[r * 2 * nz + k] = mu[m];
          hw[r * 2 * nz + nz + k] = lv[m];
        }
        ex[m] = expf(lv[m]);
        const float zv = mu[m] + sqrtf(ex[m]) * e;                             // :102-103
        a.z[m][r * nz + k] = a.round_z ? round_tf32(zv) : zv;
        row_kl[m] += 1.0f + lv[m] - mu[m] * mu[m] - ex[m];                     // :335-337 (element)
      }
      float gmu[NMOD], glv[NMOD];
#pragma unroll
      for (int m = 0; m < NMOD; ++m) {
        const float w = a.weight[m] * a.inv_global_batch;
        gmu[m] = w * mu[m];
        glv[m] = w * 0.5f * (ex[m] - 1.0f);
      }
#pragma unroll
      for (int p = 0; p < NMOD; ++p) {
#pragma unroll
        for (int q = p + 1; q < NMOD; ++q) {                                    // itertools.combinations, :346
          const float d = mu[p] - mu[q];
          const float ip = expf(-lv[p]), iq = expf(-lv[q]);
          const float epq = expf(lv[p] - lv[q]), eqp = expf(lv[q] - lv[p]);
          // 0.5*(lq - lp - 1 + e^{lp-lq} + d^2 e^{-lq}) + 0.5*(lp - lq - 1 + e^{lq-lp} + d^2 e^{-lp})   :355-365
          row_assoc += 0.5f * (epq + eqp - 2.0f + d * d * (ip + iq));
          gmu[p] += a.lambda * d * (ip + iq);
          gmu[q] -= a.lambda * d * (ip + iq);
          glv[p] += a.lambda * 0.5f * (epq - eqp - d * d * ip);
          glv[q] += a.lambda * 0.5f * (eqp - epq - d * d * iq);
        }
      }
      if (a.with_grad) {
#pragma unroll
        for (int m = 0; m < NMOD; ++m) {
          a.gstat[m][r * 2 * nz + k] = gmu[m];
          a.gstat[m][r * 2 * nz + nz + k] = glv[m];
        }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < NMOD; ++m) {
    row_kl[m] *= -0.5f;
    if (a.latent_loss[m]) a.latent_loss[m][r] = row_kl[m];                   // vae_latent_losses probe, :339
  }
}

// (A fully vectorised n_z = 4 form of this row -- 128-bit loads and stores, all results held until the end -- measured
// SLOWER inside the tile kernel: 6.2 against 4.0 us per task at B = 100, 0.282 against 0.276 ms per step at 8192.)

// element (r, k) of the latent backward for modality m: (d mu, d log sigma^2) from d z and the stashed KL gradients
template <typename LOAD>
__device__ __forceinline__ void latent_bwd_elem(const LatentBwdArgs& a, int m, int64_t r, int k, float& dm, float& dl) {
  const int nz = a.n_z;
  const float e = LOAD::ld(a.eps + r * nz + k);
  const float lv = LOAD::ld(a.heads[m] + r * 2 * nz + nz + k);
  const float s = sqrtf(expf(lv));                                           // d z / d lv = eps * s / 2
  const float dz = LOAD::ld(a.dz[m] + r * nz + k);
  dm = dz + LOAD::ld(a.gstat[m] + r * 2 * nz + k);
  dl = dz * e * 0.5f * s + LOAD::ld(a.gstat[m] + r * 2 * nz + nz + k);
  if (a.round_out) { dm = round_tf32(dm); dl = round_tf32(dl); }
  a.dheads[m][r * 2 * nz + k] = dm;
  a.dheads[m][r * 2 * nz + nz + k] = dl;
}

// n_z == 4 (the reference's latent width): every operand of the row is ONE 128-bit L1-bypassing load, so nothing of the
// row waits behind a store; m = modality.  Same arithmetic as latent_bwd_elem.  REDUCE (tile-kernel task): the warp's 32 rows are summed per k and lane 0 adds the sums into bh_grad (bias gradient of
// the heads layer: column sums of d mu | d log sigma^2); rows that are not `live` contribute zeros and store nothing.
template <bool REDUCE>
__device__ __forceinline__ void latent_bwd_row4(const LatentBwdArgs& a, int m, int64_t r, bool live, float* bh_grad, int lane) {
  float vm[4] = {0.f, 0.f, 0.f, 0.f}, vl[4] = {0.f, 0.f, 0.f, 0.f};
  if (live) {
    const float4 e4 = __ldcg(reinterpret_cast<const float4*>(a.eps + r * 4));
    const float4 lv4 = __ldcg(reinterpret_cast<const float4*>(a.heads[m] + r * 8 + 4));
    const float4 dz4 = __ldcg(reinterpret_cast<const float4*>(a.dz[m] + r * 4));
    const float4 gm4 = __ldcg(reinterpret_cast<const float4*>(a.gstat[m] + r * 8));
    const float4 gl4 = __ldcg(reinterpret_cast<const float4*>(a.gstat[m] + r * 8 + 4));
    const float e[4] = {e4.x, e4.y, e4.z, e4.w}, lv[4] = {lv4.x, lv4.y, lv4.z, lv4.w}, dz[4] = {dz4.x, dz4.y, dz4.z, dz4.w};
    const float gm[4] = {gm4.x, gm4.y, gm4.z, gm4.w}, gl[4] = {gl4.x, gl4.y, gl4.z, gl4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float s = sqrtf(expf(lv[k]));                                      // d z / d lv = eps * s / 2
      vm[k] = dz[k] + gm[k];
      vl[k] = dz[k] * e[k] * 0.5f * s + gl[k];
      if (a.round_out) { vm[k] = round_tf32(vm[k]); vl[k] = round_tf32(vl[k]); }
    }
    *reinterpret_cast<float4*>(a.dheads[m] + r * 8) = make_float4(vm[0], vm[1], vm[2], vm[3]);
    *reinterpret_cast<float4*>(a.dheads[m] + r * 8 + 4) = make_float4(vl[0], vl[1], vl[2], vl[3]);
  }
  if (REDUCE && bh_grad != nullptr) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float sm = warp_sum(vm[k]), sl = warp_sum(vl[k]);
      if (lane == 0) { atomicAdd(bh_grad + k, sm); atomicAdd(bh_grad + 4 + k, sl); }
    }
  }
}

// ---- reconstruction loss of one element from the decoder's pre-activation `a` (fused into the output-layer epilogue
// of the tile kernel; tf32 mode, tolerance 2e-3: ex2 / lg2 / rcp approximations, five MUFU operations per element) ----
constexpr float kCeEps = 1e-3f;   // vae_assoc.py:322-323 (the comment there says 1e-10; the code says 1e-3)

// Bernoulli cross-entropy with the reference's clamp: x_hat = sigmoid(a), loss = -(x log(1e-3 + x_hat) +
// (1 - x) log(1e-3 + 1 - x_hat))  (:321-324); da = scale * d loss / d a, the two quotients over one reciprocal
__device__ __forceinline__ float recon_logit_binary(float a, float x, float scale, float& da) {
  const float xh = __fdividef(1.0f, 1.0f + __expf(-a));
  const float p = kCeEps + xh;
  const float q = (kCeEps + 1.0f) - xh;              // evaluation order of :323
  da = scale * ((1.0f - x) * p - x * q) * __fdividef(xh * (1.0f - xh), p * q);
  return -(x * __logf(p) + (1.0f - x) * __logf(q));
}
// Gaussian: tf.nn.l2_loss(x_hat - x) (:327-328), x_hat = a
__device__ __forceinline__ float recon_logit_gaussian(float a, float x, float scale, float& da) {
  const float d = a - x;
  da = scale * d;
  return 0.5f * d * d;
}

// sums[2m] = reconstruction sum of modality m, sums[2m+1] = its prior-KL sum, sums[8] = association-KL sum (this rank's
// rows): per-modality cost :340, total :368-371; one thread
__device__ __forceinline__ void finalize_combine(const FinalizeArgs& a, const float* sums, int advance) {
  float cost = 0.f;
  for (int m = 0; m < a.n_mod; ++m) {
    const float rec = a.binary[m] ? sums[2 * m] * a.inv_global_batch : sums[2 * m];   // :324 per-row | :328 scalar
    const float c = (rec + sums[2 * m + 1] * a.inv_global_batch) * a.weight[m];       // :340
    a.scalars[m] = c;
    a.scalars[4 + m] = sums[2 * m];
    cost += c;
  }
  a.scalars[8] = sums[8];
  cost += a.lambda * sums[8];                                                          // :369
  a.scalars[9] = cost;
  if (a.cost_slot) *a.cost_slot = cost;
  if (advance && a.step_dev) *a.step_dev += 1;
}

}  // namespace vaeassoc
