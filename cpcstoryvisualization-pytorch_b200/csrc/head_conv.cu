// img / img_seg heads of the generator (reference model.py:272-274, 298-300): conv3x3 with 3 or 1
// output channels on the 64x64 feature map, then tanh.
//
// With so few output channels a tensor-core GEMM (N padded to 16) is bound by re-reading the
// activation tile once per filter tap through L2 (measured 180-220 us per call); these kernels
// read each activation once from HBM into a shared-memory halo tile and do the arithmetic in fp32
// on the CUDA cores, which makes the heads HBM-bound (activation planes in, NCHW image out).
//
//   forward : y[n, co, h, w] = tanh(sum_{ky,kx,ci} a[n, h+ky-1, w+kx-1, ci] * w[co, ci, ky, kx]),
//             a = hi (+ lo) 16-bit operand planes, NHWC
//   backward: g = dy * (1 - y^2);
//             da[n, h, w, ci] (+)= sum_{ky,kx,co} g[n, co, h-ky+1, w-kx+1] * w[co, ci, ky, kx]
//             dw[co, ci, ky, kx] = sum_{n,h,w} g[n, co, h, w] * a_hi[n, h+ky-1, w+kx-1, ci]
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.h"

namespace cpcsv {

constexpr int kHeadTileH = 8;     // output rows per CTA
constexpr int kHeadTileW = 64;    // output columns per CTA
constexpr int kHeadChunk = 16;    // input channels staged per pass (32 B = one sector per pixel and plane)
constexpr int kHeadPitch = 20;    // floats per staged pixel (16 + 4: conflict-free float4 access)
constexpr int kHeadThreads = 128; // thread = (column, strip of 4 rows)
constexpr int kHeadPix = (kHeadTileH + 2) * (kHeadTileW + 2);
constexpr int kHeadIters = (kHeadPix + kHeadThreads - 1) / kHeadThreads;   // halo pixels per thread

__device__ __forceinline__ float cvt16(uint16_t v, int dtype) {
  if (dtype == 1) return __uint_as_float(static_cast<uint32_t>(v) << 16);
  return __half2float(__ushort_as_half(v));
}

__device__ __forceinline__ void unpack8(const uint4& q, int dtype, float* f) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = cvt16(static_cast<uint16_t>(w[i] & 0xffffu), dtype);
    f[2 * i + 1] = cvt16(static_cast<uint16_t>(w[i] >> 16), dtype);
  }
}

// Halo staging, software pipelined: `fetch` issues the global loads of channels [c0, c0 + 16) of
// the (8 + 2) x (64 + 2) halo tile into registers (all of a thread's pixels back to back),
// `commit` converts them to fp32 (hi + lo) and writes them to shared memory.  The kernel fetches
// chunk k + 1 before it multiplies chunk k, so the HBM latency is hidden behind the arithmetic.
struct HaloRegs {
  uint4 h[kHeadIters][2];
  uint4 l[kHeadIters][2];
  uint32_t ok;
};

__device__ __forceinline__ void halo_fetch(HaloRegs& R, const uint16_t* __restrict__ hi,
                                           const uint16_t* __restrict__ lo, int n, int H, int W, int C,
                                           int y0, int x0, int c0) {
  R.ok = 0;
#pragma unroll
  for (int b = 0; b < kHeadIters; ++b) {
    const int idx = threadIdx.x + b * kHeadThreads;
    const int r = idx / (kHeadTileW + 2), cx = idx - r * (kHeadTileW + 2);
    const int gy = y0 - 1 + r, gx = x0 - 1 + cx;
    if (idx < kHeadPix && gy >= 0 && gy < H && gx >= 0 && gx < W) {
      R.ok |= 1u << b;
      const int64_t off = ((static_cast<int64_t>(n) * H + gy) * W + gx) * C + c0;
      const uint4* ph = reinterpret_cast<const uint4*>(hi + off);
      R.h[b][0] = __ldg(ph);
      R.h[b][1] = __ldg(ph + 1);
      if (lo) {
        const uint4* pl = reinterpret_cast<const uint4*>(lo + off);
        R.l[b][0] = __ldg(pl);
        R.l[b][1] = __ldg(pl + 1);
      }
    }
  }
}

__device__ __forceinline__ void halo_commit(const HaloRegs& R, bool two, int dtype, float* a_s) {
#pragma unroll
  for (int b = 0; b < kHeadIters; ++b) {
    const int idx = threadIdx.x + b * kHeadThreads;
    if (idx >= kHeadPix) break;
    float f[16];
    if ((R.ok >> b) & 1u) {
      unpack8(R.h[b][0], dtype, f);
      unpack8(R.h[b][1], dtype, f + 8);
      if (two) {
        float g[16];
        unpack8(R.l[b][0], dtype, g);
        unpack8(R.l[b][1], dtype, g + 8);
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] += g[i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = 0.f;
    }
    float4* dst = reinterpret_cast<float4*>(a_s + idx * kHeadPitch);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
  }
}

template <int CO>
__global__ void __launch_bounds__(kHeadThreads, 3)
head_conv_tanh_fwd_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, int dtype,
                          int N, int H, int W, int C, const float* __restrict__ w,
                          float* __restrict__ y) {
  extern __shared__ float smem_f[];
  float* a_s = smem_f;                                                       // [10][66][20]
  float* w_s = smem_f + (kHeadTileH + 2) * (kHeadTileW + 2) * kHeadPitch;    // [9][CO][16]
  const int tiles_x = W / kHeadTileW, tiles_y = H / kHeadTileH;
  int t = blockIdx.x;
  const int tx = t % tiles_x;
  t /= tiles_x;
  const int ty = t % tiles_y;
  const int n = t / tiles_y;
  const int x0 = tx * kHeadTileW, y0 = ty * kHeadTileH;
  const int x = threadIdx.x & (kHeadTileW - 1);
  const int strip = threadIdx.x / kHeadTileW;  // rows strip*4 .. strip*4+3 of the tile
  float acc[4][CO];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int co = 0; co < CO; ++co) acc[i][co] = 0.f;

  HaloRegs R;
  halo_fetch(R, hi, lo, n, H, W, C, y0, x0, 0);
  for (int c0 = 0; c0 < C; c0 += kHeadChunk) {
    halo_commit(R, lo != nullptr, dtype, a_s);
    for (int idx = threadIdx.x; idx < 9 * CO * kHeadChunk; idx += kHeadThreads) {
      const int c = idx % kHeadChunk;
      const int q = idx / kHeadChunk;
      const int co = q % CO, tap = q / CO;
      w_s[idx] = __ldg(w + (static_cast<int64_t>(co) * C + c0 + c) * 9 + tap);
    }
    __syncthreads();
    if (c0 + kHeadChunk < C) halo_fetch(R, hi, lo, n, H, W, C, y0, x0, c0 + kHeadChunk);
#pragma unroll
    for (int c4 = 0; c4 < kHeadChunk / 4; ++c4) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        // the three taps of this filter column stay in registers (9 float4 for CO = 3)
        float4 wr[3][CO];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int co = 0; co < CO; ++co)
            wr[ky][co] = *reinterpret_cast<const float4*>(w_s + ((ky * 3 + dx) * CO + co) * kHeadChunk + 4 * c4);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const float4 v = *reinterpret_cast<const float4*>(
              a_s + ((strip * 4 + r) * (kHeadTileW + 2) + x + dx) * kHeadPitch + 4 * c4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int ky = r - i;
            if (ky >= 0 && ky < 3) {
#pragma unroll
              for (int co = 0; co < CO; ++co) {
                const float4 ww = wr[ky][co];
                acc[i][co] = fmaf(v.x, ww.x, acc[i][co]);
                acc[i][co] = fmaf(v.y, ww.y, acc[i][co]);
                acc[i][co] = fmaf(v.z, ww.z, acc[i][co]);
                acc[i][co] = fmaf(v.w, ww.w, acc[i][co]);
              }
            }
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gy = y0 + strip * 4 + i;
#pragma unroll
    for (int co = 0; co < CO; ++co)
      y[((static_cast<int64_t>(n) * CO + co) * H + gy) * W + x0 + x] = tanhf(acc[i][co]);
  }
}

template <int CO>
static int launch_head_fwd(unsigned grid, size_t smem, cudaStream_t stream, const uint16_t* hi,
                           const uint16_t* lo, int dtype, int N, int H, int W, int C, const float* w,
                           float* y) {
  static cudaError_t attr = cudaFuncSetAttribute(
      head_conv_tanh_fwd_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  if (attr != cudaSuccess)
    return fail(static_cast<int>(attr), "head_conv_tanh_fwd: cudaFuncSetAttribute: %s",
                cudaGetErrorString(attr));
  head_conv_tanh_fwd_kernel<CO><<<grid, kHeadThreads, smem, stream>>>(hi, lo, dtype, N, H, W, C, w, y);
  return 0;
}

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

extern "C" int cpcsv_head_conv_tanh_fwd(const void* a_hi, const void* a_lo, int32_t dtype, int32_t N,
                                        int32_t H, int32_t W, int32_t C, const float* w, int32_t Co,
                                        float* y, cpcsv_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CPCSV_REQUIRE(a_hi && w && y && N > 0, "head_conv_tanh_fwd: args");
  CPCSV_REQUIRE(dtype == 0 || dtype == 1, "head_conv_tanh_fwd: dtype %d", dtype);
  CPCSV_REQUIRE(H % kHeadTileH == 0 && W % kHeadTileW == 0 && C % kHeadChunk == 0,
                "head_conv_tanh_fwd: needs H %% 8 == 0, W %% 64 == 0, C %% 16 == 0 (got %d x %d x %d)", H,
                W, C);
  CPCSV_REQUIRE(Co >= 1 && Co <= 3, "head_conv_tanh_fwd: Co %d (1..3)", Co);
  CPCSV_REQUIRE((reinterpret_cast<uintptr_t>(a_hi) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a_lo) & 15) == 0,
                "head_conv_tanh_fwd: operand planes must be 16 B aligned");
  const size_t smem =
      sizeof(float) * ((kHeadTileH + 2) * (kHeadTileW + 2) * kHeadPitch + 9 * Co * kHeadChunk);
  const unsigned grid = static_cast<unsigned>(N * (H / kHeadTileH) * (W / kHeadTileW));
  const uint16_t* hi = static_cast<const uint16_t*>(a_hi);
  const uint16_t* lo = static_cast<const uint16_t*>(a_lo);
  int rc;
  switch (Co) {
    case 1: rc = launch_head_fwd<1>(grid, smem, stream, hi, lo, dtype, N, H, W, C, w, y); break;
    case 2: rc = launch_head_fwd<2>(grid, smem, stream, hi, lo, dtype, N, H, W, C, w, y); break;
    default: rc = launch_head_fwd<3>(grid, smem, stream, hi, lo, dtype, N, H, W, C, w, y); break;
  }
  if (rc) return rc;
  return launched("head_conv_tanh_fwd");
}
