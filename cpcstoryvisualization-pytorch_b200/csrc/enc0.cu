// First layer of the discriminators' encode_img (reference model.py:498-500, 540-542, 582-584):
// Conv2d(C, ndf, 4, 2, 1, bias=False) on the 3- or 1-channel 64x64 image + LeakyReLU(0.2), straight to the
// NHWC 16-bit operand planes of the next layer's GEMM.
//
// K = 16 C = 48 or 16: as an implicit GEMM the layer needed an im2col launch (K padded to 64), a GEMM whose
// tiles are all epilogue, and an activation / re-pack launch over its fp32 output -- three trips through HBM
// for 1.1 GFLOP.  Here one CTA computes 2 output rows x 32 output columns x all (<= 128) output channels in fp32
// on the CUDA cores from a shared-memory input patch and the whole (transposed) weight matrix, and writes the
// activated hi / lo planes once: 4.4 MB read, 47 MB written per 90-image call.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.h"

namespace cpcsv {
namespace {

__device__ __forceinline__ void split16(float v, int dtype, uint16_t& hi, uint16_t& lo) {
  if (dtype == 1) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
  } else {
    __half h = __float2half_rn(v);
    __half l = __float2half_rn(v - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
  }
}

constexpr int kWS = 132;   // weight row pitch in shared memory (floats): 16-byte rows, 4-way conflicts on the fill
constexpr int kPW = 68;    // patch row pitch

// grid: N * (OH / 2) * (OW / 32) CTAs of 256 threads; thread = (pixel group pg = tid / 16, channel group cg = tid % 16):
// pixels pg + 16 j (j = 0..3) of the 2 x 32 tile, channels 8 cg .. 8 cg + 7.
template <int C>
__global__ void __launch_bounds__(256)
enc0_lrelu_fwd_kernel(const float* __restrict__ x, int64_t sn, int64_t sc, int64_t sh, int64_t sw, int H, int W,
                      const float* __restrict__ w, int Co, const float* __restrict__ alpha, float slope,
                      uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int ldp, int dtype) {
  constexpr int K = C * 16;
  __shared__ __align__(16) float w_s[K * kWS];
  __shared__ float patch[C][6][kPW];
  const int OH = H / 2, OW = W / 2;
  const int tiles_w = OW / 32, tiles_h = OH / 2;
  const int bx = blockIdx.x % tiles_w;
  const int by = (blockIdx.x / tiles_w) % tiles_h;
  const int n = blockIdx.x / (tiles_w * tiles_h);
  const int tid = threadIdx.x;
  // weights [Co][C][4][4] -> w_s[k][co], zero beyond Co
  for (int idx = tid; idx < 128 * K; idx += 256) {
    const int co = idx / K, k = idx - co * K;
    w_s[k * kWS + co] = co < Co ? __ldg(w + idx) : 0.f;
  }
  const int ih0 = 4 * by - 1, iw0 = 64 * bx - 1;
  for (int idx = tid; idx < C * 6 * 66; idx += 256) {
    const int col = idx % 66;
    const int r = (idx / 66) % 6;
    const int c = idx / (66 * 6);
    const int ih = ih0 + r, iw = iw0 + col;
    float v = 0.f;
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __ldg(x + n * sn + c * sc + ih * sh + iw * sw);
    patch[c][r][col] = v;
  }
  __syncthreads();
  const int cg = tid & 15, pg = tid >> 4;
  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 4; ++kx) {
        const int k = c * 16 + ky * 4 + kx;
        const float4 w0 = *reinterpret_cast<const float4*>(&w_s[k * kWS + cg * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&w_s[k * kWS + cg * 8 + 4]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int p = pg + 16 * j;
          const float xv = patch[c][2 * (p >> 5) + ky][2 * (p & 31) + kx];
          acc[j][0] = fmaf(xv, w0.x, acc[j][0]);
          acc[j][1] = fmaf(xv, w0.y, acc[j][1]);
          acc[j][2] = fmaf(xv, w0.z, acc[j][2]);
          acc[j][3] = fmaf(xv, w0.w, acc[j][3]);
          acc[j][4] = fmaf(xv, w1.x, acc[j][4]);
          acc[j][5] = fmaf(xv, w1.y, acc[j][5]);
          acc[j][6] = fmaf(xv, w1.z, acc[j][6]);
          acc[j][7] = fmaf(xv, w1.w, acc[j][7]);
        }
      }
    }
  }
  if (cg * 8 >= ldp) return;
  const float a = alpha ? __ldg(alpha) : 1.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = pg + 16 * j;
    const int64_t row = (static_cast<int64_t>(n) * OH + 2 * by + (p >> 5)) * OW + 32 * bx + (p & 31);
    uint16_t hv[8], lv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = acc[j][i] * a;
      v = v > 0.f ? v : slope * v;
      split16(v, dtype, hv[i], lv[i]);
    }
    uint4 ph, pl;
    ph.x = hv[0] | (static_cast<uint32_t>(hv[1]) << 16);
    ph.y = hv[2] | (static_cast<uint32_t>(hv[3]) << 16);
    ph.z = hv[4] | (static_cast<uint32_t>(hv[5]) << 16);
    ph.w = hv[6] | (static_cast<uint32_t>(hv[7]) << 16);
    *reinterpret_cast<uint4*>(hi + row * ldp + cg * 8) = ph;
    if (lo) {
      pl.x = lv[0] | (static_cast<uint32_t>(lv[1]) << 16);
      pl.y = lv[2] | (static_cast<uint32_t>(lv[3]) << 16);
      pl.z = lv[4] | (static_cast<uint32_t>(lv[5]) << 16);
      pl.w = lv[6] | (static_cast<uint32_t>(lv[7]) << 16);
      *reinterpret_cast<uint4*>(lo + row * ldp + cg * 8) = pl;
    }
  }
}

// dz (bf16) = dy * (a > 0 ? 1 : slope), a = the activated value's hi plane (LeakyReLU keeps the sign)
__global__ void lrelu_bwd16_kernel(const float* __restrict__ dy, const uint16_t* __restrict__ a_hi, int64_t n8,
                                   float slope, uint16_t* __restrict__ dz) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(dy) + 2 * i);
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(dy) + 2 * i + 1);
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(a_hi) + i);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      // positive <=> sign bit clear and magnitude bits not all zero (both 16-bit formats)
      const uint32_t lo16 = aw[q] & 0xffffu, hi16 = aw[q] >> 16;
      const bool p0 = (lo16 & 0x8000u) == 0u && (lo16 & 0x7fffu) != 0u;
      const bool p1 = (hi16 & 0x8000u) == 0u && (hi16 & 0x7fffu) != 0u;
      const float v0 = p0 ? g[2 * q] : slope * g[2 * q];
      const float v1 = p1 ? g[2 * q + 1] : slope * g[2 * q + 1];
      o[q] = __bfloat16_as_ushort(__float2bfloat16_rn(v0)) |
             (static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(v1))) << 16);
    }
    reinterpret_cast<uint4*>(dz)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace
}  // namespace cpcsv

using namespace cpcsv;

extern "C" int cpcsv_enc0_lrelu_fwd(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, int64_t sn,
                                    int64_t sc, int64_t sh, int64_t sw, const float* w, int32_t Co,
                                    const float* alpha, float slope, void* hi, void* lo, int32_t ldp,
                                    int32_t dtype, cpcsv_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CPCSV_REQUIRE(x && w && hi && N > 0 && (C == 1 || C == 3), "enc0_lrelu_fwd: 1- or 3-channel images");
  CPCSV_REQUIRE(H > 0 && H % 4 == 0 && W > 0 && W % 64 == 0, "enc0_lrelu_fwd: H %% 4, W %% 64 (got %d x %d)", H, W);
  CPCSV_REQUIRE(Co > 0 && Co <= ldp && ldp <= 128 && ldp % 8 == 0, "enc0_lrelu_fwd: Co %d, channel pitch %d", Co, ldp);
  const unsigned grid = static_cast<unsigned>(static_cast<int64_t>(N) * (H / 4) * (W / 64));
  uint16_t* h16 = static_cast<uint16_t*>(hi);
  uint16_t* l16 = static_cast<uint16_t*>(lo);
  if (C == 3)
    enc0_lrelu_fwd_kernel<3><<<grid, 256, 0, stream>>>(x, sn, sc, sh, sw, H, W, w, Co, alpha, slope, h16, l16, ldp, dtype);
  else
    enc0_lrelu_fwd_kernel<1><<<grid, 256, 0, stream>>>(x, sn, sc, sh, sw, H, W, w, Co, alpha, slope, h16, l16, ldp, dtype);
  return launched("enc0_lrelu_fwd");
}

extern "C" int cpcsv_lrelu_bwd16(const float* dy, const void* a_hi, int64_t count, float slope, void* dz,
                                 cpcsv_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CPCSV_REQUIRE(dy && a_hi && dz && count > 0 && count % 8 == 0, "lrelu_bwd16: count %% 8");
  const int64_t n8 = count / 8;
  int64_t blocks = ceil_div(n8, 256);
  const int64_t cap = static_cast<int64_t>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  lrelu_bwd16_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(dy, static_cast<const uint16_t*>(a_hi), n8,
                                                                         slope, static_cast<uint16_t*>(dz));
  return launched("lrelu_bwd16");
}
