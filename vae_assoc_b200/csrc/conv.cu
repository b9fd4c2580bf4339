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
#include "common.cuh"
#include "kernels.h"

namespace vaeassoc {

namespace {

// out[(b,oy,ox), (ky,kx,c)] = x[b, oy*s+ky-pb, ox*s+kx-pb, c]  (0 outside); one thread per (row, ky, kx, c4-chunk)
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

inline int grid_cap(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = 32 * kNumSMs;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

void launch_im2col(const Im2colArgs& a, cudaStream_t s) {
  im2col_kernel<<<grid_cap((int64_t)a.B * a.OH * a.OW * a.k * a.k * a.C), 256, 0, s>>>(a);
}

void launch_col2im(const Col2imArgs& a, cudaStream_t s) {
  col2im_kernel<<<grid_cap((int64_t)a.B * a.H * a.W * a.C), 256, 0, s>>>(a);
}

}  // namespace vaeassoc
