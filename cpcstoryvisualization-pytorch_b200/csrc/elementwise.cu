// HBM-bound kernels around the tcgen05 GEMMs: BatchNorm statistics / apply / backward fused
// with the activation and with the split of fp32 activations into 16-bit hi/lo operand planes,
// plus the layout kernels (NCHW <-> NHWC, im2col of the 1/3-channel images, head tanh).
// All are coalesced along the channel (innermost NHWC) dimension and vectorised 4-8 wide.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.h"

namespace cpcsv {

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == 1) return v > 0.f ? v : 0.f;
  if (act == 2) return v > 0.f ? v : 0.2f * v;
  return v;
}
__device__ __forceinline__ float act_grad(float pre, int act) {
  if (act == 1) return pre > 0.f ? 1.f : 0.f;
  if (act == 2) return pre > 0.f ? 1.f : 0.2f;
  return 1.f;
}

// split v into a 16-bit hi part and the 16-bit rounding of the remainder
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
__device__ __forceinline__ uint16_t to16(float v, int dtype) {
  return dtype == 1 ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
}

static inline int grid_for(int64_t work_items, int threads, int max_blocks_per_sm = 8) {
  int64_t blocks = ceil_div(work_items, threads);
  int64_t cap = static_cast<int64_t>(num_sms()) * max_blocks_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ------------------------------------------------------------------------- pack_nchw
__global__ void pack_nchw_kernel(const float* __restrict__ x, int N, int C, int H, int W, int64_t sn,
                                 int64_t sc, int64_t sh, int64_t sw, const float* __restrict__ bcast,
                                 int Cb, int64_t ldb, uint16_t* __restrict__ hi,
                                 uint16_t* __restrict__ lo, int Cpad, int dtype) {
  const int64_t total = static_cast<int64_t>(N) * H * W * Cpad;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cpad);
    int64_t p = i / Cpad;
    const int w = static_cast<int>(p % W);
    p /= W;
    const int h = static_cast<int>(p % H);
    const int n = static_cast<int>(p / H);
    float v = 0.f;
    if (c < C) v = x[n * sn + c * sc + h * sh + w * sw];
    else if (c < C + Cb) v = bcast[n * ldb + (c - C)];
    uint16_t hv, lv;
    split16(v, dtype, hv, lv);
    hi[i] = hv;
    if (lo) lo[i] = lv;
  }
}

// ------------------------------------------------------------------------- im2col (small C)
__global__ void im2col_small_kernel(const float* __restrict__ x, int N, int C, int H, int W,
                                    int64_t sn, int64_t sc, int64_t sh, int64_t sw, int k, int s,
                                    int p, int OH, int OW, uint16_t* __restrict__ hi,
                                    uint16_t* __restrict__ lo, int ldp, int dtype) {
  const int64_t total = static_cast<int64_t>(N) * OH * OW * ldp;
  const int kk = k * k * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % ldp);
    int64_t q = i / ldp;
    const int ow = static_cast<int>(q % OW);
    q /= OW;
    const int oh = static_cast<int>(q % OH);
    const int n = static_cast<int>(q / OH);
    float v = 0.f;
    if (j < kk) {
      const int tap = j / C, c = j - tap * C;
      const int ky = tap / k, kx = tap - ky * k;
      const int ih = oh * s - p + ky, iw = ow * s - p + kx;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = x[n * sn + c * sc + ih * sh + iw * sw];
    }
    uint16_t hv, lv;
    split16(v, dtype, hv, lv);
    hi[i] = hv;
    if (lo) lo[i] = lv;
  }
}

__global__ void col2im_small_kernel(const float* __restrict__ dcol, int64_t ldc, int N, int C, int H,
                                    int W, int k, int s, int p, int OH, int OW,
                                    float* __restrict__ dx) {
  const int64_t total = static_cast<int64_t>(N) * C * H * W;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int w = static_cast<int>(i % W);
    int64_t q = i / W;
    const int h = static_cast<int>(q % H);
    q /= H;
    const int c = static_cast<int>(q % C);
    const int n = static_cast<int>(q / C);
    float acc = 0.f;
    for (int ky = 0; ky < k; ++ky) {
      const int t = h + p - ky;
      if (t < 0 || t % s) continue;
      const int oh = t / s;
      if (oh >= OH) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int u = w + p - kx;
        if (u < 0 || u % s) continue;
        const int ow = u / s;
        if (ow >= OW) continue;
        acc += dcol[(static_cast<int64_t>(n) * OH * OW + oh * OW + ow) * ldc + (ky * k + kx) * C + c];
      }
    }
    dx[i] = acc;
  }
}

// ------------------------------------------------------------------------- image output
// [C, H, W] fp32 in [-1, 1] (any strides) -> uint8 [H, W, C]: the conversion the reference does on the host
// for every generated frame / sheet (miscc/utils.py:230-235 images_to_numpy); on the device the read-back is
// a quarter of the bytes and already in the layout PIL / the PNG writer wants.
__global__ void images_to_u8_kernel(const float* __restrict__ x, int C, int H, int W, int64_t sc, int64_t sh,
                                    int64_t sw, uint8_t* __restrict__ out) {
  const int64_t total = static_cast<int64_t>(H) * W * C;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int64_t p = i / C;
    const int w = static_cast<int>(p % W), h = static_cast<int>(p / W);
    float v = x[c * sc + h * sh + w * sw];
    v = fminf(fmaxf(v, -1.f), 1.f);
    // (v + 1) / 2 * 255 truncated, exactly the host formula
    out[i] = static_cast<uint8_t>((v + 1.0f) / 2.0f * 255.0f);
  }
}

// ------------------------------------------------------------------------- heads
__global__ void tanh_to_nchw_kernel(const float* __restrict__ z, int64_t ldz, int N, int C, int H,
                                    int W, float* __restrict__ y) {
  const int64_t total = static_cast<int64_t>(N) * C * H * W;
  const int64_t hw = static_cast<int64_t>(H) * W;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t pix_in_img = i % hw;
    const int64_t q = i / hw;
    const int c = static_cast<int>(q % C);
    const int64_t n = q / C;
    y[i] = tanhf(z[(n * hw + pix_in_img) * ldz + c]);
  }
}

__global__ void tanh_bwd_im2col_kernel(const float* __restrict__ dy, int64_t sn, int64_t sc,
                                       int64_t sh, int64_t sw, const float* __restrict__ y, int N,
                                       int C, int H, int W, uint16_t* __restrict__ col, int ldp,
                                       int dtype) {
  const int64_t total = static_cast<int64_t>(N) * H * W * ldp;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % ldp);
    int64_t q = i / ldp;
    const int w = static_cast<int>(q % W);
    q /= W;
    const int h = static_cast<int>(q % H);
    const int n = static_cast<int>(q / H);
    float v = 0.f;
    if (j < 9 * C) {
      const int tap = j / C, c = j - tap * C;
      const int ky = tap / 3, kx = tap - ky * 3;
      const int ih = h - (ky - 1), iw = w - (kx - 1);
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
        const float yy = y[((static_cast<int64_t>(n) * C + c) * H + ih) * W + iw];
        v = dy[n * sn + c * sc + ih * sh + iw * sw] * (1.f - yy * yy);
      }
    }
    col[i] = to16(v, dtype);
  }
}

// Row-staged variant (C <= 4, W <= 128): one block per image row.  dz = dy * (1 - y^2) of the
// three source rows is computed once into shared memory with coalesced NCHW reads; the 64-wide
// im2col rows are then written contiguously.
__global__ void __launch_bounds__(256)
tanh_bwd_im2col_rows_kernel(const float* __restrict__ dy, int64_t sn, int64_t sc, int64_t sh,
                            int64_t sw, const float* __restrict__ y, int C, int H, int W,
                            uint16_t* __restrict__ col, int ldp, int dtype) {
  __shared__ float dz[3][4][130];   // [source row h-1..h+1][channel][w + 1], zero borders
  const int h = blockIdx.x % H;
  const int n = blockIdx.x / H;
  for (int i = threadIdx.x; i < 3 * C * (W + 2); i += 256) {
    const int wq = i % (W + 2);
    const int c = (i / (W + 2)) % C;
    const int rr = i / ((W + 2) * C);
    const int ih = h - 1 + rr, iw = wq - 1;
    float v = 0.f;
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      const float yy = y[((static_cast<int64_t>(n) * C + c) * H + ih) * W + iw];
      v = dy[n * sn + c * sc + ih * sh + iw * sw] * (1.f - yy * yy);
    }
    dz[rr][c][wq] = v;
  }
  __syncthreads();
  uint16_t* out = col + (static_cast<int64_t>(n) * H + h) * W * ldp;
  for (int i = threadIdx.x; i < W * ldp; i += 256) {
    const int j = i % ldp, w = i / ldp;
    float v = 0.f;
    if (j < 9 * C) {
      const int tap = j / C, c = j - tap * C;
      const int ky = tap / 3, kx = tap - ky * 3;
      // source pixel (h - (ky-1), w - (kx-1)) -> staged row 2 - ky, column w - kx + 2
      v = dz[2 - ky][c][w - kx + 2];
    }
    out[i] = to16(v, dtype);
  }
}

// ------------------------------------------------------------------------- weight packing
__global__ void pack_matrix_kernel(const float* __restrict__ w, int64_t rows_out, int64_t cols_out,
                                   int64_t cols_valid, int64_t ld_r, int64_t ld_c,
                                   const int32_t* __restrict__ row_map,
                                   const int32_t* __restrict__ col_map, uint16_t* __restrict__ hi,
                                   uint16_t* __restrict__ lo, int64_t ldo, int dtype) {
  const int64_t total = rows_out * cols_out;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols_out, c = i - r * cols_out;
    float v = 0.f;
    if (c < cols_valid) {
      const int64_t rs = row_map ? row_map[r] : r;
      const int64_t cs = col_map ? col_map[c] : c;
      if (rs >= 0 && cs >= 0) v = w[rs * ld_r + cs * ld_c];
    }
    uint16_t hv, lv;
    split16(v, dtype, hv, lv);
    hi[r * ldo + c] = hv;
    if (lo) lo[r * ldo + c] = lv;
  }
}

__global__ void scatter_rows_kernel(const float* __restrict__ src, int64_t ld_src,
                                    const int32_t* __restrict__ row_map, float* __restrict__ dst,
                                    int64_t ld_dst, int64_t rows, int64_t cols) {
  const int64_t total = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const int64_t rd = row_map ? row_map[r] : r;
    if (rd >= 0) dst[rd * ld_dst + c] = src[r * ld_src + c];
  }
}

// taps of the 3x3 kernel merged into sub-pixel tap i of phase a:
//   a=0: i=0 -> {0},   i=1 -> {1,2};   a=1: i=0 -> {0,1}, i=1 -> {2}
__device__ __forceinline__ void merged_range(int a, int i, int& lo, int& hi) {
  if (a == 0) { lo = i == 0 ? 0 : 1; hi = i == 0 ? 0 : 2; }
  else        { lo = i == 0 ? 0 : 2; hi = i == 0 ? 1 : 2; }
}

// Tiled variant: a block owns a 16 (co) x 32 (ci) tile of the weight for ALL taps.  The source
// w[co][ci][t] is read as contiguous runs of 32*T floats per co, staged in shared memory, and
// written out per tap with the threads running along the destination's contiguous dimension.
constexpr int kPackCo = 8, kPackCi = 16;

__global__ void __launch_bounds__(256)
pack_conv_weight_tiled_kernel(const float* __restrict__ w, int Cout, int Cin, int T, int kind,
                              int rows_pad, int cols_pad, uint16_t* __restrict__ hi,
                              uint16_t* __restrict__ lo, int dtype) {
  // 8 (co) x 16 (ci) x T tile = 8.2 KB of shared memory: small enough to be resident next to a
  // tensor-core GEMM block of another stream (the re-layout runs on the prefetch stream)
  __shared__ float s[kPackCo][kPackCi * 16 + 1];
  const int co0 = blockIdx.y * kPackCo, ci0 = blockIdx.x * kPackCi;
  const bool transposed = (kind == 1 || kind == 3);
  const int ntap = kind >= 2 ? 16 : T;
  // load: kPackCo rows (co) of kPackCi*T contiguous floats
  const int run = kPackCi * T;
  for (int i = threadIdx.x; i < kPackCo * run; i += 256) {
    const int r = i / run, j = i - r * run;
    const int co = co0 + r, ci = ci0 + j / T;
    s[r][j] = (co < Cout && ci < Cin) ? w[(static_cast<int64_t>(co) * Cin + ci0) * T + j] : 0.f;
  }
  __syncthreads();
  // store: every thread emits 8 consecutive elements (one 16-byte store per plane) along the
  // output's contiguous dimension: ci for [tap][co][ci], co for [tap][ci][co]
  constexpr int kVec = kPackCo * kPackCi / 8;   // 16-byte output vectors per tap
  for (int i = threadIdx.x; i < ntap * kVec; i += 256) {
    const int tap = i / kVec, v = i % kVec;
    int r0, c0, dr, dc;      // first (co, ci) element inside the tile and the step along the vector
    if (!transposed) { c0 = (v % (kPackCi / 8)) * 8; r0 = v / (kPackCi / 8); dr = 0; dc = 1; }
    else             { r0 = (v % (kPackCo / 8)) * 8; c0 = v / (kPackCo / 8); dr = 1; dc = 0; }
    int y0 = 0, y1 = 0, x0 = 0, x1 = 0;
    if (kind >= 2) {
      const int a = tap >> 3, b = (tap >> 2) & 1, ti = (tap >> 1) & 1, tj = tap & 1;
      merged_range(a, ti, y0, y1);
      merged_range(b, tj, x0, x1);
    }
    uint16_t hv[8], lv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int r = r0 + e * dr, c = c0 + e * dc;
      float v;
      if (kind < 2) {
        v = s[r][c * T + tap];
      } else {
        v = 0.f;
        for (int ky = y0; ky <= y1; ++ky)
          for (int kx = x0; kx <= x1; ++kx) v += s[r][c * 9 + ky * 3 + kx];
      }
      split16(v, dtype, hv[e], lv[e]);
    }
    const int row = transposed ? ci0 + c0 : co0 + r0;
    const int col = transposed ? co0 + r0 : ci0 + c0;
    if (row < rows_pad && col < cols_pad) {
      const int64_t o = (static_cast<int64_t>(tap) * rows_pad + row) * cols_pad + col;
      uint4 hq, lq;
      hq.x = hv[0] | (static_cast<uint32_t>(hv[1]) << 16); hq.y = hv[2] | (static_cast<uint32_t>(hv[3]) << 16);
      hq.z = hv[4] | (static_cast<uint32_t>(hv[5]) << 16); hq.w = hv[6] | (static_cast<uint32_t>(hv[7]) << 16);
      *reinterpret_cast<uint4*>(hi + o) = hq;
      if (lo) {
        lq.x = lv[0] | (static_cast<uint32_t>(lv[1]) << 16); lq.y = lv[2] | (static_cast<uint32_t>(lv[3]) << 16);
        lq.z = lv[4] | (static_cast<uint32_t>(lv[5]) << 16); lq.w = lv[6] | (static_cast<uint32_t>(lv[7]) << 16);
        *reinterpret_cast<uint4*>(lo + o) = lq;
      }
    }
  }
}

__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int kh, int kw,
                                        int kind, int rows_pad, int cols_pad,
                                        uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                        int dtype) {
  const int ntap = (kind >= 2) ? 16 : kh * kw;
  const int64_t total = static_cast<int64_t>(ntap) * rows_pad * cols_pad;
  const bool transposed = (kind == 1 || kind == 3);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cols_pad);
    int64_t q = i / cols_pad;
    const int r = static_cast<int>(q % rows_pad);
    const int tap = static_cast<int>(q / rows_pad);
    const int co = transposed ? c : r;
    const int ci = transposed ? r : c;
    float v = 0.f;
    if (co < Cout && ci < Cin) {
      const float* base = w + (static_cast<int64_t>(co) * Cin + ci) * kh * kw;
      if (kind < 2) {
        v = base[tap];
      } else {
        const int a = tap >> 3, b = (tap >> 2) & 1, ti = (tap >> 1) & 1, tj = tap & 1;
        int y0, y1, x0, x1;
        merged_range(a, ti, y0, y1);
        merged_range(b, tj, x0, x1);
        for (int ky = y0; ky <= y1; ++ky)
          for (int kx = x0; kx <= x1; ++kx) v += base[ky * 3 + kx];
      }
    }
    uint16_t hv, lv;
    split16(v, dtype, hv, lv);
    hi[i] = hv;
    if (lo) lo[i] = lv;
  }
}

// Coalesced variant for the untransposed layouts dwt[tap][co][ci]: one block = one `co` and 128
// consecutive `ci`; reads are contiguous in ci, the T = kh*kw values of every (co, ci) are
// transposed through shared memory so the writes to dw[co][ci][t] are contiguous too.
// wdot / dot (optional): also accumulates sum(dw .* wdot) into *dot -- the scalar the spectral-norm
// backward needs (cpcsv_spectral_bwd_apply), without another pass over the gradient.
__global__ void __launch_bounds__(128)
unpack_conv_wgrad_rows_kernel(const float* __restrict__ dwt, int64_t mat_stride, int64_t ldc, int Cout,
                              int Cin, int T, int kind, const float* __restrict__ alpha,
                              float* __restrict__ dw, const float* __restrict__ wdot,
                              float* __restrict__ dot) {
  __shared__ float s[128 * 16];
  const int co = blockIdx.y;
  const int ci0 = blockIdx.x * 128;
  const int ci = ci0 + threadIdx.x;
  const float a_ = alpha ? *alpha : 1.f;
  const int tin = kind == 2 ? 16 : T;
  float v[16];
#pragma unroll
  for (int t = 0; t < 16; ++t)
    v[t] = (t < tin && ci < Cin) ? dwt[t * mat_stride + static_cast<int64_t>(co) * ldc + ci] : 0.f;
  if (kind == 2) {
    // tap index = a*8 + b*4 + i*2 + j; 3x3 tap (ky,kx) collects the phases that merged it
    float o[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float acc = 0.f;
#pragma unroll
        for (int tap = 0; tap < 16; ++tap) {
          const int a = tap >> 3, b = (tap >> 2) & 1, ti = (tap >> 1) & 1, tj = tap & 1;
          int y0, y1, x0, x1;
          merged_range(a, ti, y0, y1);
          merged_range(b, tj, x0, x1);
          if (ky >= y0 && ky <= y1 && kx >= x0 && kx <= x1) acc += v[tap];
        }
        o[ky * 3 + kx] = acc;
      }
#pragma unroll
    for (int t = 0; t < 9; ++t) s[threadIdx.x * T + t] = o[t] * a_;
  } else {
#pragma unroll
    for (int t = 0; t < 16; ++t)
      if (t < T) s[threadIdx.x * T + t] = v[t] * a_;
  }
  __syncthreads();
  const int nvalid = min(128, Cin - ci0);
  const int64_t base = (static_cast<int64_t>(co) * Cin + ci0) * T;
  float* dst = dw + base;
  float part = 0.f;
  for (int j = threadIdx.x; j < nvalid * T; j += 128) {
    const float g = s[j];
    dst[j] = g;
    if (wdot) part = fmaf(g, wdot[base + j], part);
  }
  if (wdot) {
    __shared__ float red[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(dot, red[0] + red[1] + red[2] + red[3]);
  }
}

__global__ void unpack_conv_wgrad_kernel(const float* __restrict__ dwt, int64_t mat_stride,
                                         int64_t ldc, int Cout, int Cin, int kh, int kw, int kind,
                                         const float* __restrict__ alpha, float* __restrict__ dw) {
  const int64_t total = static_cast<int64_t>(Cout) * Cin * kh * kw;
  const bool transposed = (kind == 1 || kind == 3);
  const float a_ = alpha ? *alpha : 1.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i % (kh * kw));
    int64_t q = i / (kh * kw);
    const int ci = static_cast<int>(q % Cin);
    const int co = static_cast<int>(q / Cin);
    const int64_t elem = transposed ? static_cast<int64_t>(ci) * ldc + co
                                    : static_cast<int64_t>(co) * ldc + ci;
    float v = 0.f;
    if (kind < 2) {
      v = dwt[t * mat_stride + elem];
    } else {
      const int ky = t / 3, kx = t - ky * 3;
      for (int a = 0; a < 2; ++a)
        for (int ti = 0; ti < 2; ++ti) {
          int y0, y1;
          merged_range(a, ti, y0, y1);
          if (ky < y0 || ky > y1) continue;
          for (int b = 0; b < 2; ++b)
            for (int tj = 0; tj < 2; ++tj) {
              int x0, x1;
              merged_range(b, tj, x0, x1);
              if (kx < x0 || kx > x1) continue;
              const int tap = (a << 3) | (b << 2) | (ti << 1) | tj;
              v += dwt[tap * mat_stride + elem];
            }
        }
    }
    dw[i] = v * a_;
  }
}

}  // namespace cpcsv

using namespace cpcsv;

#define STREAM(s) static_cast<cudaStream_t>(s)

// (the BatchNorm entry points live in bn.cu)

extern "C" int cpcsv_images_to_u8(const float* x, int32_t C, int32_t H, int32_t W, int64_t sc, int64_t sh,
                                  int64_t sw, uint8_t* out, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(x && out && C > 0 && H > 0 && W > 0, "images_to_u8: args");
  const int64_t work = static_cast<int64_t>(C) * H * W;
  images_to_u8_kernel<<<grid_for(work, 256), 256, 0, STREAM(stream)>>>(x, C, H, W, sc, sh, sw, out);
  return launched("images_to_u8");
}

extern "C" int cpcsv_pack_nchw(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, int64_t sn,
                               int64_t sc, int64_t sh, int64_t sw, const float* bcast, int32_t Cb,
                               int64_t ldb, void* hi, void* lo, int32_t Cpad, int32_t dtype,
                               cpcsv_stream_t stream) {
  CPCSV_REQUIRE(x && hi && N > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C + Cb, "pack_nchw: args");
  const int64_t work = static_cast<int64_t>(N) * H * W * Cpad;
  pack_nchw_kernel<<<grid_for(work, 256), 256, 0, STREAM(stream)>>>(
      x, N, C, H, W, sn, sc, sh, sw, bcast, Cb, ldb, static_cast<uint16_t*>(hi),
      static_cast<uint16_t*>(lo), Cpad, dtype);
  return launched("pack_nchw");
}

extern "C" int cpcsv_im2col_small(const float* x, int32_t N, int32_t C, int32_t H, int32_t W,
                                  int64_t sn, int64_t sc, int64_t sh, int64_t sw, int32_t k, int32_t s,
                                  int32_t p, void* hi, void* lo, int32_t ldp, int32_t dtype,
                                  cpcsv_stream_t stream) {
  CPCSV_REQUIRE(x && hi && N > 0 && C > 0 && k * k * C <= ldp, "im2col_small: args");
  const int OH = (H + 2 * p - k) / s + 1, OW = (W + 2 * p - k) / s + 1;
  const int64_t work = static_cast<int64_t>(N) * OH * OW * ldp;
  im2col_small_kernel<<<grid_for(work, 256), 256, 0, STREAM(stream)>>>(
      x, N, C, H, W, sn, sc, sh, sw, k, s, p, OH, OW, static_cast<uint16_t*>(hi),
      static_cast<uint16_t*>(lo), ldp, dtype);
  return launched("im2col_small");
}

extern "C" int cpcsv_col2im_small(const float* dcol, int64_t ldc, int32_t N, int32_t C, int32_t H,
                                  int32_t W, int32_t k, int32_t s, int32_t p, float* dx,
                                  cpcsv_stream_t stream) {
  CPCSV_REQUIRE(dcol && dx && N > 0 && C > 0, "col2im_small: args");
  const int OH = (H + 2 * p - k) / s + 1, OW = (W + 2 * p - k) / s + 1;
  const int64_t work = static_cast<int64_t>(N) * C * H * W;
  col2im_small_kernel<<<grid_for(work, 256), 256, 0, STREAM(stream)>>>(dcol, ldc, N, C, H, W, k, s, p,
                                                                       OH, OW, dx);
  return launched("col2im_small");
}

extern "C" int cpcsv_tanh_to_nchw(const float* z, int64_t ldz, int32_t N, int32_t C, int32_t H,
                                  int32_t W, float* y, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(z && y && N > 0 && C > 0, "tanh_to_nchw: args");
  const int64_t work = static_cast<int64_t>(N) * C * H * W;
  tanh_to_nchw_kernel<<<grid_for(work, 256), 256, 0, STREAM(stream)>>>(z, ldz, N, C, H, W, y);
  return launched("tanh_to_nchw");
}

extern "C" int cpcsv_tanh_bwd_im2col(const float* dy, int64_t sn, int64_t sc, int64_t sh, int64_t sw,
                                     const float* y, int32_t N, int32_t C, int32_t H, int32_t W,
                                     void* col, int32_t ldp, int32_t dtype, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(dy && y && col && 9 * C <= ldp, "tanh_bwd_im2col: args");
  const int64_t work = static_cast<int64_t>(N) * H * W * ldp;
  if (C <= 4 && W <= 128) {
    tanh_bwd_im2col_rows_kernel<<<static_cast<unsigned>(N * H), 256, 0, STREAM(stream)>>>(
        dy, sn, sc, sh, sw, y, C, H, W, static_cast<uint16_t*>(col), ldp, dtype);
    return launched("tanh_bwd_im2col");
  }
  tanh_bwd_im2col_kernel<<<grid_for(work, 256), 256, 0, STREAM(stream)>>>(
      dy, sn, sc, sh, sw, y, N, C, H, W, static_cast<uint16_t*>(col), ldp, dtype);
  return launched("tanh_bwd_im2col");
}

extern "C" int cpcsv_pack_matrix(const float* w, int64_t rows_out, int64_t cols_out,
                                 int64_t cols_valid, int64_t ld_r, int64_t ld_c,
                                 const int32_t* row_map, const int32_t* col_map, void* hi, void* lo,
                                 int64_t ldo, int32_t dtype, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(w && hi && rows_out > 0 && cols_out > 0 && ldo >= cols_out, "pack_matrix: args");
  pack_matrix_kernel<<<grid_for(rows_out * cols_out, 256), 256, 0, STREAM(stream)>>>(
      w, rows_out, cols_out, cols_valid, ld_r, ld_c, row_map, col_map, static_cast<uint16_t*>(hi),
      static_cast<uint16_t*>(lo), ldo, dtype);
  return launched("pack_matrix");
}

extern "C" int cpcsv_scatter_rows_f32(const float* src, int64_t ld_src, const int32_t* row_map,
                                      float* dst, int64_t ld_dst, int64_t rows, int64_t cols,
                                      cpcsv_stream_t stream) {
  CPCSV_REQUIRE(src && dst && rows > 0 && cols > 0, "scatter_rows_f32: args");
  scatter_rows_kernel<<<grid_for(rows * cols, 256), 256, 0, STREAM(stream)>>>(src, ld_src, row_map, dst,
                                                                             ld_dst, rows, cols);
  return launched("scatter_rows_f32");
}

extern "C" int cpcsv_pack_conv_weight(const float* w, int32_t Cout, int32_t Cin, int32_t kh,
                                      int32_t kw, int32_t kind, int32_t rows_pad, int32_t cols_pad,
                                      void* hi, void* lo, int32_t dtype, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(w && hi && kind >= 0 && kind <= 3, "pack_conv_weight: args");
  CPCSV_REQUIRE(kind < 2 || (kh == 3 && kw == 3), "pack_conv_weight: sub-pixel merge needs 3x3");
  const bool tr = (kind == 1 || kind == 3);
  CPCSV_REQUIRE(rows_pad >= (tr ? Cin : Cout) && cols_pad >= (tr ? Cout : Cin),
                "pack_conv_weight: padding smaller than the matrix");
  const int ntap = kind >= 2 ? 16 : kh * kw;
  const int64_t work = static_cast<int64_t>(ntap) * rows_pad * cols_pad;
  if (kh * kw <= 16 && cols_pad % 8 == 0) {
    // tiles cover the PADDED index space so the zero padding is written too
    const int co_ext = tr ? cols_pad : rows_pad, ci_ext = tr ? rows_pad : cols_pad;
    dim3 grid(static_cast<unsigned>(ceil_div(ci_ext, kPackCi)),
              static_cast<unsigned>(ceil_div(co_ext, kPackCo)));
    pack_conv_weight_tiled_kernel<<<grid, 256, 0, STREAM(stream)>>>(
        w, Cout, Cin, kh * kw, kind, rows_pad, cols_pad, static_cast<uint16_t*>(hi),
        static_cast<uint16_t*>(lo), dtype);
    return launched("pack_conv_weight");
  }
  pack_conv_weight_kernel<<<grid_for(work, 256), 256, 0, STREAM(stream)>>>(
      w, Cout, Cin, kh, kw, kind, rows_pad, cols_pad, static_cast<uint16_t*>(hi),
      static_cast<uint16_t*>(lo), dtype);
  return launched("pack_conv_weight");
}

extern "C" int cpcsv_unpack_conv_wgrad(const float* dwt, int64_t mat_stride, int64_t ldc,
                                       int32_t Cout, int32_t Cin, int32_t kh, int32_t kw,
                                       int32_t kind, const float* alpha, float* dw,
                                       cpcsv_stream_t stream) {
  CPCSV_REQUIRE(dwt && dw && kind >= 0 && kind <= 3, "unpack_conv_wgrad: args");
  CPCSV_REQUIRE(kind < 2 || (kh == 3 && kw == 3), "unpack_conv_wgrad: sub-pixel merge needs 3x3");
  const int64_t work = static_cast<int64_t>(Cout) * Cin * kh * kw;
  if ((kind == 0 || kind == 2) && kh * kw <= 16) {
    dim3 grid(static_cast<unsigned>(ceil_div(Cin, 128)), static_cast<unsigned>(Cout));
    unpack_conv_wgrad_rows_kernel<<<grid, 128, 0, STREAM(stream)>>>(dwt, mat_stride, ldc, Cout, Cin,
                                                                   kh * kw, kind, alpha, dw, nullptr, nullptr);
    return launched("unpack_conv_wgrad");
  }
  unpack_conv_wgrad_kernel<<<grid_for(work, 256), 256, 0, STREAM(stream)>>>(
      dwt, mat_stride, ldc, Cout, Cin, kh, kw, kind, alpha, dw);
  return launched("unpack_conv_wgrad");
}

extern "C" int cpcsv_unpack_conv_wgrad_dot(const float* dwt, int64_t mat_stride, int64_t ldc,
                                           int32_t Cout, int32_t Cin, int32_t kh, int32_t kw,
                                           int32_t kind, const float* alpha, float* dw, const float* w,
                                           float* dot, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(dwt && dw && w && dot && (kind == 0 || kind == 2) && kh * kw <= 16,
                "unpack_conv_wgrad_dot: args (untransposed kinds, <= 16 taps)");
  CPCSV_REQUIRE(kind < 2 || (kh == 3 && kw == 3), "unpack_conv_wgrad_dot: sub-pixel merge needs 3x3");
  dim3 grid(static_cast<unsigned>(ceil_div(Cin, 128)), static_cast<unsigned>(Cout));
  unpack_conv_wgrad_rows_kernel<<<grid, 128, 0, STREAM(stream)>>>(dwt, mat_stride, ldc, Cout, Cin, kh * kw,
                                                                 kind, alpha, dw, w, dot);
  return launched("unpack_conv_wgrad_dot");
}
