// img / img_seg heads of the generator (reference model.py:272-274, 298-300): conv3x3 with 3 or 1
// output channels on the 64x64 feature map, then tanh.
//
// With so few output channels an implicit GEMM over the 9 taps (N padded to 16) re-reads the activation
// tile once per tap through L2 (180-220 us per call), and a direct CUDA-core kernel from shared-memory
// halo tiles is bound by its L1-tag-limited staging loads and the fp32 FMAs (137-178 us; round 1).
#include "common.h"

namespace cpcsv {

// Heads as ONE pixel-major GEMM plus a gather: Z[p, tap*Co + co] = sum_c a[p, c] * w[co, c, tap] on the
// tensor cores (the activation is read once, N = 9*Co padded to 16 / 32), then
//   y[n, co, h, w] = tanh(sum_{ky,kx} Z[(n, h+ky-1, w+kx-1), (ky*3+kx)*Co + co])   (zero outside the image).
// Z (128 B per pixel) has just been written and is L2-resident; a warp walks 32 consecutive columns, so
// every Z line it touches is used by three taps of its neighbours.
template <int CO>
__global__ void __launch_bounds__(256)
head_gather_tanh_kernel(const float* __restrict__ z, int64_t ldz, int N, int H, int W,
                        float* __restrict__ y) {
  const int64_t total = static_cast<int64_t>(N) * H * W;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < total;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(p % W);
    const int h = static_cast<int>((p / W) % H);
    const int64_t n = p / (static_cast<int64_t>(W) * H);
    float acc[CO];
#pragma unroll
    for (int co = 0; co < CO; ++co) acc[co] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int hh = h + ky - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ww = w + kx - 1;
        if (ww < 0 || ww >= W) continue;
        const float* src = z + ((n * H + hh) * W + ww) * ldz + (ky * 3 + kx) * CO;
#pragma unroll
        for (int co = 0; co < CO; ++co) acc[co] += __ldg(src + co);
      }
    }
#pragma unroll
    for (int co = 0; co < CO; ++co) y[((n * CO + co) * H + h) * W + w] = tanhf(acc[co]);
  }
}

}  // namespace cpcsv

using namespace cpcsv;

extern "C" int cpcsv_head_gather_tanh(const float* z, int64_t ldz, int32_t N, int32_t H, int32_t W, int32_t Co,
                                      float* y, cpcsv_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CPCSV_REQUIRE(z && y && N > 0 && H > 0 && W > 0 && Co >= 1 && Co <= 3 && ldz >= 9 * Co,
                "head_gather_tanh: args");
  const int64_t total = static_cast<int64_t>(N) * H * W;
  int64_t blocks = ceil_div(total, 256);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  const unsigned grid = static_cast<unsigned>(blocks);
  switch (Co) {
    case 1: head_gather_tanh_kernel<1><<<grid, 256, 0, stream>>>(z, ldz, N, H, W, y); break;
    case 2: head_gather_tanh_kernel<2><<<grid, 256, 0, stream>>>(z, ldz, N, H, W, y); break;
    default: head_gather_tanh_kernel<3><<<grid, 256, 0, stream>>>(z, ldz, N, H, W, y); break;
  }
  return launched("head_gather_tanh");
}
