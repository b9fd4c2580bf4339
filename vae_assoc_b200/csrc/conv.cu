// Data-movement kernels of the hidden_conv=True image modality.
//
// Every conv / transposed-conv layer of the reference is expressed as  im2col -> dense contraction -> col2im  so that
// the arithmetic runs on the library's GEMM kernels (tcgen05 for the 400/800-deep layers):
//   conv_2d        vae_assoc.py:480-489  (tf.nn.conv2d, NHWC, filter [k,k,Cin,Cout] = matrix [k*k*Cin, Cout])
//                  y = im2col(x) . W                      dW = im2col(x)^T . dy        dx = col2im(dy . W^T)
//   deconv2d       deconv.py:29-128     (tf.nn.conv2d_transpose, filter [k,k,Cout,Cin] = matrix [k*k*Cout, Cin])
//                  out = act(col2im(y . W^T) + bias)      dW = im2col(dout)^T . y      dy = im2col(dout) . W
// TensorFlow padding: SAME  -> out = ceil(in/s), pad_before = max((out-1)s+k-in, 0)/2 ; VALID -> no padding.
// conv2d_transpose(SAME, s) is the gradient of that conv: out[oy] += in[iy] w[ky] with oy = iy*s + ky - pad_before
// (one pixel off torch's padding=2/output_padding=1 convention).
// Both kernels are pure gathers (no atomics, deterministic), HBM-bound: algorithmic bytes = 4 * (input + output).
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace vaeassoc {

namespace {

// out[(b,oy,ox), (ky,kx,c)] = x[b, oy*s+ky-pb, ox*s+kx-pb, c]  (0 outside); scalar form: one thread per element
__global__ void __launch_bounds__(256) im2col_kernel(Im2colArgs a) {
  const int kk = a.k * a.k;
  const int64_t rows = (int64_t)a.B * a.OH * a.OW;
  const int64_t total = rows * kk * a.C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % a.C);
    int64_t t = i / a.C;
    const int kx = (int)(t % a.k); t /= a.k;
    const int ky = (int)(t % a.k); t /= a.k;       // t = output row (b, oy, ox)
    const int ox = (int)(t % a.OW);
    int64_t u = t / a.OW;
    const int oy = (int)(u % a.OH);
    const int64_t b = u / a.OH;
    const int iy = oy * a.s + ky - a.pb, ix = ox * a.s + kx - a.pb;
    float v = 0.f;
    if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) v = a.x[((b * a.H + iy) * a.W + ix) * a.C + c];
    a.out[t * a.ldo + (ky * a.k + kx) * a.C + c] = v;
  }
}

// C == 1 (the 28 x 28 x 1 image itself): one warp per patch row, one lane per (ky, kx) tap -> one contiguous row store
__global__ void __launch_bounds__(256) im2col_c1_kernel(Im2colArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kk = a.k * a.k;
  const int ky = lane / a.k, kx = lane - ky * a.k;
  const int64_t rows = (int64_t)a.B * a.OH * a.OW;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += (int64_t)gridDim.x * 8) {
    const int ox = (int)(r % a.OW);
    const int64_t u = r / a.OW;
    const int oy = (int)(u % a.OH);
    const int64_t b = u / a.OH;
    if (lane < kk) {
      const int iy = oy * a.s + ky - a.pb, ix = ox * a.s + kx - a.pb;
      float v = 0.f;
      if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) v = __ldg(a.x + (b * a.H + iy) * a.W + ix);
      a.out[r * a.ldo + lane] = v;
    }
  }
}

// C % 4 == 0: one warp per patch row, lanes stride over its k*k*C/4 float4 elements -> 128-bit loads (contiguous along
// (kx, c) in the input) and 128-bit stores (contiguous along the output row); the (ky, kx, c) decomposition of an
// element index comes from a per-block table, the (b, oy, ox) decomposition is done once per row
constexpr int kIm2colMaxElems = 25 * 16;           // float4 elements of a patch row handled by the table (k <= 5, C <= 64)
__global__ void __launch_bounds__(256) im2col_v4_kernel(Im2colArgs a) {
  __shared__ int s_tab[kIm2colMaxElems];           // ky | kx << 8 | (c4 * 4) << 16
  const int c4n = a.C >> 2, ne = a.k * a.k * c4n;
  for (int e = threadIdx.x; e < ne; e += blockDim.x) {
    const int kidx = e / c4n, c4 = e - kidx * c4n;
    const int ky = kidx / a.k, kx = kidx - ky * a.k;
    s_tab[e] = ky | (kx << 8) | ((c4 * 4) << 16);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)a.B * a.OH * a.OW;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < rows; r += (int64_t)gridDim.x * 8) {
    const int ox = (int)(r % a.OW);
    const int64_t u = r / a.OW;
    const int oy = (int)(u % a.OH);
    const int64_t b = u / a.OH;
    const int y0 = oy * a.s - a.pb, x0 = ox * a.s - a.pb;
    const float* __restrict__ img = a.x + b * a.H * a.W * a.C;
    float4* __restrict__ dst = reinterpret_cast<float4*>(a.out + r * a.ldo);
    for (int e = lane; e < ne; e += 32) {
      const int tb = s_tab[e];
      const int iy = y0 + (tb & 0xff), ix = x0 + ((tb >> 8) & 0xff), c = tb >> 16;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W)
        v = __ldg(reinterpret_cast<const float4*>(img + ((int64_t)iy * a.W + ix) * a.C + c));
      dst[e] = v;
    }
  }
}

// out[b,y,x,c] = act( sum_{ky,kx} cols[(b,iy,ix), (ky,kx,c)] + bias[c] ),  iy = (y + pb - ky)/s when divisible and in range
__global__ void __launch_bounds__(256) col2im_kernel(Col2imArgs a) {
  const int64_t total = (int64_t)a.B * a.H * a.W * a.C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % a.C);
    int64_t t = i / a.C;
    const int x = (int)(t % a.W); t /= a.W;
    const int y = (int)(t % a.H);
    const int64_t b = t / a.H;
    float acc = 0.f;
    for (int ky = 0; ky < a.k; ++ky) {
      const int ty = y + a.pb - ky;
      if (ty < 0 || ty % a.s) continue;
      const int iy = ty / a.s;
      if (iy >= a.h) continue;
      for (int kx = 0; kx < a.k; ++kx) {
        const int tx = x + a.pb - kx;
        if (tx < 0 || tx % a.s) continue;
        const int ix = tx / a.s;
        if (ix >= a.w) continue;
        acc += a.cols[((b * a.h + iy) * a.w + ix) * a.ldc + (ky * a.k + kx) * a.C + c];
      }
    }
    if (a.bias) acc += __ldg(a.bias + c);
    acc = apply_act(a.act, acc);
    if (a.round_out) acc = round_tf32(acc);
    a.out[i] = acc;
  }
}

// C % 4 == 0: one thread per float4 of the output (same tap order as the scalar form: ky, then kx); S = stride (1 or 2)
template <int S>
__global__ void __launch_bounds__(256) col2im_v4_kernel(Col2imArgs a) {
  const int c4n = a.C >> 2;
  const int64_t total = (int64_t)a.B * a.H * a.W * c4n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    int64_t t = i / c4n;
    const int x = (int)(t % a.W); t /= a.W;
    const int y = (int)(t % a.H);
    const int64_t b = t / a.H;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* __restrict__ base = a.cols + b * a.h * a.w * a.ldc + c;
    for (int ky = 0; ky < a.k; ++ky) {
      const int ty = y + a.pb - ky;
      if (ty < 0 || (ty % S)) continue;
      const int iy = ty / S;
      if (iy >= a.h) continue;
      for (int kx = 0; kx < a.k; ++kx) {
        const int tx = x + a.pb - kx;
        if (tx < 0 || (tx % S)) continue;
        const int ix = tx / S;
        if (ix >= a.w) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(base + ((int64_t)iy * a.w + ix) * a.ldc + (ky * a.k + kx) * a.C));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (a.bias) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + c));
      acc.x += bv.x; acc.y += bv.y; acc.z += bv.z; acc.w += bv.w;
    }
    acc.x = apply_act(a.act, acc.x); acc.y = apply_act(a.act, acc.y); acc.z = apply_act(a.act, acc.z); acc.w = apply_act(a.act, acc.w);
    if (a.round_out) { acc.x = round_tf32(acc.x); acc.y = round_tf32(acc.y); acc.z = round_tf32(acc.z); acc.w = round_tf32(acc.w); }
    *reinterpret_cast<float4*>(a.out + ((b * a.H + y) * a.W + x) * a.C + c) = acc;
  }
}

// C == 1 (the last deconv writes the 28 x 28 x 1 image): one warp per output image row (b, y), lane = x; the ky / iy
// tests are warp-uniform, the kx / ix tests per lane; same tap order as the scalar form
template <int S>
__global__ void __launch_bounds__(256) col2im_c1_kernel(Col2imArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t nrows = (int64_t)a.B * a.H;
  const float bias = a.bias ? __ldg(a.bias) : 0.f;
  for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < nrows; r += (int64_t)gridDim.x * 8) {
    const int y = (int)(r % a.H);
    const int64_t b = r / a.H;
    for (int x = lane; x < a.W; x += 32) {
      float acc = 0.f;
      for (int ky = 0; ky < a.k; ++ky) {
        const int ty = y + a.pb - ky;
        if (ty < 0 || (ty % S)) continue;
        const int iy = ty / S;
        if (iy >= a.h) continue;
        const float* __restrict__ row = a.cols + ((b * a.h + iy) * a.w) * a.ldc + ky * a.k;
        for (int kx = 0; kx < a.k; ++kx) {
          const int tx = x + a.pb - kx;
          if (tx < 0 || (tx % S)) continue;
          const int ix = tx / S;
          if (ix >= a.w) continue;
          acc += __ldg(row + (int64_t)ix * a.ldc + kx);
        }
      }
      acc = apply_act(a.act, acc + bias);
      if (a.round_out) acc = round_tf32(acc);
      a.out[(b * a.H + y) * a.W + x] = acc;
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline int grid_cap(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = 32 * kNumSMs;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

void launch_im2col(const Im2colArgs& a, cudaStream_t s) {
  if (a.C % 4 == 0 && a.ldo % 4 == 0 && a.k * a.k * (a.C / 4) <= kIm2colMaxElems && a.k < 256 && aligned16(a.x) && aligned16(a.out)) {
    const int64_t rows = (int64_t)a.B * a.OH * a.OW;
    const int64_t blocks = (rows + 7) / 8;
    im2col_v4_kernel<<<(int)std::min<int64_t>(std::max<int64_t>(blocks, 1), 64 * kNumSMs), 256, 0, s>>>(a);
    return;
  }
  if (a.C == 1 && a.k * a.k <= 32) {
    const int64_t blocks = ((int64_t)a.B * a.OH * a.OW + 7) / 8;
    im2col_c1_kernel<<<(int)std::min<int64_t>(std::max<int64_t>(blocks, 1), 64 * kNumSMs), 256, 0, s>>>(a);
    return;
  }
  im2col_kernel<<<grid_cap((int64_t)a.B * a.OH * a.OW * a.k * a.k * a.C), 256, 0, s>>>(a);
}

void launch_col2im(const Col2imArgs& a, cudaStream_t s) {
  if (a.C % 4 == 0 && a.ldc % 4 == 0 && (a.s == 1 || a.s == 2) && aligned16(a.cols) && aligned16(a.out) &&
      (a.bias == nullptr || aligned16(a.bias))) {
    const int grid = grid_cap((int64_t)a.B * a.H * a.W * (a.C / 4));
    if (a.s == 1) col2im_v4_kernel<1><<<grid, 256, 0, s>>>(a);
    else col2im_v4_kernel<2><<<grid, 256, 0, s>>>(a);
    return;
  }
  if (a.C == 1 && (a.s == 1 || a.s == 2)) {
    const int64_t blocks = ((int64_t)a.B * a.H + 7) / 8;
    const int grid = (int)std::min<int64_t>(std::max<int64_t>(blocks, 1), 64 * kNumSMs);
    if (a.s == 1) col2im_c1_kernel<1><<<grid, 256, 0, s>>>(a);
    else col2im_c1_kernel<2><<<grid, 256, 0, s>>>(a);
    return;
  }
  col2im_kernel<<<grid_cap((int64_t)a.B * a.H * a.W * a.C), 256, 0, s>>>(a);
}

}  // namespace vaeassoc
