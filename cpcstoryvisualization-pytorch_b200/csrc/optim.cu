// Adam (reference trainer.py:212-220: optim.Adam(lr, betas=(0.5, 0.999)); steps at l.345-346, 416)
// as hand-written multi-tensor kernels, FUSED with the re-layout of the updated weights into the
// 16-bit operand planes the tcgen05 GEMMs read.  One pass over p / g / m / v per optimiser step
// writes p, m, v and every packed plane of the parameter, so no separate pack kernel (and no
// second read of the fp32 weight) runs between an optimiser step and the next forward pass.
//
// Arithmetic = torch.optim.Adam (fp32, no weight decay, no amsgrad):
//   m += (g - m) * (1 - beta1);  v = beta2 * v + (1 - beta2) * g * g
//   p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// The step count lives on the device (graph replays advance it); `cpcsv_adam_tick` turns it into
// the two bias-correction factors once per optimiser step.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.h"

namespace cpcsv {
namespace {

struct Hyper {
  const float* lr;   // device scalar
  const float* bc;   // device [2]: 1 / (1 - beta1^t), 1 / sqrt(1 - beta2^t)
  float beta1, beta2, omb1, omb2, eps;   // omb = 1 - beta, rounded from the double difference
};

__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, float step_size,
                                             float inv_sqrt_bc2, const Hyper& h) {
  m = fmaf(h.omb1, g - m, m);
  v = h.beta2 * v + h.omb2 * g * g;
  const float denom = sqrtf(v) * inv_sqrt_bc2 + h.eps;
  return p - step_size * (m / denom);
}

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

__global__ void adam_tick_kernel(float* step, double beta1, double beta2, float* bc) {
  const float t = *step + 1.f;
  *step = t;
  bc[0] = static_cast<float>(1.0 / (1.0 - pow(beta1, static_cast<double>(t))));
  bc[1] = static_cast<float>(1.0 / sqrt(1.0 - pow(beta2, static_cast<double>(t))));
}

// ------------------------------------------------------------------------- generic multi-tensor
constexpr int kMaxTensors = 48;
constexpr int kChunk = 4096;   // elements per block

struct TensorTable {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  int64_t n[kMaxTensors];
  int32_t block0[kMaxTensors + 1];
  int32_t count;
};

__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ TensorTable T, const Hyper h) {
  int t = 0;
  while (t + 1 < T.count && static_cast<int>(blockIdx.x) >= T.block0[t + 1]) ++t;
  const int64_t base = static_cast<int64_t>(blockIdx.x - T.block0[t]) * kChunk;
  const int64_t n = T.n[t];
  float* __restrict__ p = T.p[t];
  const float* __restrict__ g = T.g[t];
  float* __restrict__ m = T.m[t];
  float* __restrict__ v = T.v[t];
  const float step_size = __ldg(h.lr) * __ldg(h.bc);
  const float isb2 = __ldg(h.bc + 1);
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                     reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
#pragma unroll
  for (int u = 0; u < kChunk / (256 * 4); ++u) {
    const int64_t i = base + (static_cast<int64_t>(u) * 256 + threadIdx.x) * 4;
    if (i >= n) break;
    if (vec && i + 4 <= n) {
      float4 pv = *reinterpret_cast<float4*>(p + i);
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      pv.x = adam_update(pv.x, gv.x, mv.x, vv.x, step_size, isb2, h);
      pv.y = adam_update(pv.y, gv.y, mv.y, vv.y, step_size, isb2, h);
      pv.z = adam_update(pv.z, gv.z, mv.z, vv.z, step_size, isb2, h);
      pv.w = adam_update(pv.w, gv.w, mv.w, vv.w, step_size, isb2, h);
      *reinterpret_cast<float4*>(p + i) = pv;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (int64_t j = i; j < n && j < i + 4; ++j) {
        float mm = m[j], vv = v[j];
        p[j] = adam_update(p[j], g[j], mm, vv, step_size, isb2, h);
        m[j] = mm;
        v[j] = vv;
      }
    }
  }
}

// ------------------------------------------------------------------------- conv weight + planes
struct PlaneSet {
  cpcsv_plane_t pl[CPCSV_MAX_PLANES];
  int32_t count;
};

// taps of the 3x3 kernel merged into sub-pixel tap i of phase a:
//   a=0: i=0 -> {0},   i=1 -> {1,2};   a=1: i=0 -> {0,1}, i=1 -> {2}
__device__ __forceinline__ void merged_range(int a, int i, int& lo, int& hi) {
  if (a == 0) { lo = i == 0 ? 0 : 1; hi = i == 0 ? 0 : 2; }
  else        { lo = i == 0 ? 0 : 2; hi = i == 0 ? 1 : 2; }
}

constexpr int kTile = 16;   // (co) x (ci) tile of a block

// A block owns a 16 (co) x 16 (ci) tile of w[Cout][Cin][T] for all T taps (T = 9 or 16, a template
// parameter so the index arithmetic is constant divisions).  Every thread first loads its T elements of
// p / g / m / v (contiguous runs of 16*T floats per co: coalesced, 4*T independent loads in flight),
// updates them, writes p / m / v back and stages the new weights in shared memory; the tile is then
// written out once per requested plane, 16 bytes per thread and 32-byte runs along the plane's
// contiguous dimension (ci for [tap][co][ci], co for the transposed [tap][ci][co]).
// kMerged (3x3 kernels of the up-blocks): all planes hold the 16 sub-pixel merged taps; they are formed
// ONCE per (co, ci) pair in a second shared-memory tile and every plane is then a plain copy + split.
template <int T, bool kMerged>
__global__ void __launch_bounds__(256)
adam_pack_conv_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                      float* __restrict__ v, int Cout, int Cin, const Hyper h,
                      const __grid_constant__ PlaneSet S) {
  constexpr int kRun = kTile * T;
  constexpr int kTaps = kMerged ? 16 : T;          // taps of the emitted planes
  __shared__ float s[kTile][kRun + 1];
  __shared__ float s2[kMerged ? kTile : 1][kMerged ? kTile * 17 + 1 : 1];   // 17: conflict-free per-pair writes
  const int co0 = blockIdx.y * kTile, ci0 = blockIdx.x * kTile;
  if (g) {
    const float step_size = __ldg(h.lr) * __ldg(h.bc);
    const float isb2 = __ldg(h.bc + 1);
    constexpr int kU = 8;   // elements per thread in flight (4 arrays each)
#pragma unroll
    for (int u0 = 0; u0 < T; u0 += kU) {
      float pv[kU], gv[kU], mv[kU], vv[kU];
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        const int u = u0 + k;
        if (u < T) {
          const int i = threadIdx.x + 256 * u;
          const int r = i / kRun, j = i - r * kRun;
          const int co = co0 + r, ci = ci0 + j / T;
          const bool ok = co < Cout && ci < Cin;
          const int64_t o = (static_cast<int64_t>(co) * Cin + ci0) * T + j;
          pv[k] = ok ? p[o] : 0.f;
          gv[k] = ok ? g[o] : 0.f;
          mv[k] = ok ? m[o] : 0.f;
          vv[k] = ok ? v[o] : 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        const int u = u0 + k;
        if (u < T) {
          const int i = threadIdx.x + 256 * u;
          const int r = i / kRun, j = i - r * kRun;
          const int co = co0 + r, ci = ci0 + j / T;
          if (co < Cout && ci < Cin) {
            const int64_t o = (static_cast<int64_t>(co) * Cin + ci0) * T + j;
            pv[k] = adam_update(pv[k], gv[k], mv[k], vv[k], step_size, isb2, h);
            p[o] = pv[k];
            m[o] = mv[k];
            v[o] = vv[k];
          }
          s[r][j] = pv[k];
        }
      }
    }
  } else {
#pragma unroll
    for (int u = 0; u < T; ++u) {
      const int i = threadIdx.x + 256 * u;
      const int r = i / kRun, j = i - r * kRun;
      const int co = co0 + r, ci = ci0 + j / T;
      s[r][j] = (co < Cout && ci < Cin) ? p[(static_cast<int64_t>(co) * Cin + ci0) * T + j] : 0.f;
    }
  }
  __syncthreads();
  if (kMerged) {
    // thread = one (co, ci) pair of the tile: 3x3 -> 4 phases x 2x2 merged taps
    const int r = threadIdx.x >> 4, c = threadIdx.x & 15;
    float w[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) w[t] = s[r][c * 9 + t];
#pragma unroll
    for (int tap = 0; tap < 16; ++tap) {
      const int a = tap >> 3, b = (tap >> 2) & 1, ti = (tap >> 1) & 1, tj = tap & 1;
      int y0, y1, x0, x1;
      merged_range(a, ti, y0, y1);
      merged_range(b, tj, x0, x1);
      float val = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
          if (ky >= y0 && ky <= y1 && kx >= x0 && kx <= x1) val += w[ky * 3 + kx];
      s2[r][c * 17 + tap] = val;
    }
    __syncthreads();
  }
  constexpr int kVec = kTile * kTile / 8;   // 16-byte output vectors per tap
  for (int q = 0; q < S.count; ++q) {
    const cpcsv_plane_t& P = S.pl[q];
    const int dtype = P.dtype;
    const bool transposed = (P.kind == 1 || P.kind == 3);
    uint16_t* __restrict__ hi = static_cast<uint16_t*>(P.hi);
    uint16_t* __restrict__ lo = static_cast<uint16_t*>(P.lo);
    for (int i = threadIdx.x; i < kTaps * kVec; i += 256) {
      const int tap = i / kVec, vi = i % kVec;
      // two consecutive threads write the two 16-byte halves of one 32-byte run
      const int minor = (vi & 1) * 8, major = vi >> 1;
      const int r0 = transposed ? minor : major;     // co inside the tile
      const int c0 = transposed ? major : minor;     // ci inside the tile
      const int dr = transposed ? 1 : 0, dc = transposed ? 0 : 1;
      const int row = transposed ? ci0 + c0 : co0 + r0;
      const int col = transposed ? co0 + r0 : ci0 + c0;
      if (row >= P.rows_pad || col >= P.cols_pad) continue;
      uint16_t hv[8], lv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int r = r0 + e * dr, c = c0 + e * dc;
        const float val = kMerged ? s2[r][c * 17 + tap] : s[r][c * T + tap];
        split16(val, dtype, hv[e], lv[e]);
      }
      const int64_t o = (static_cast<int64_t>(tap) * P.rows_pad + row) * P.cols_pad + col;
      uint4 hq;
      hq.x = hv[0] | (static_cast<uint32_t>(hv[1]) << 16); hq.y = hv[2] | (static_cast<uint32_t>(hv[3]) << 16);
      hq.z = hv[4] | (static_cast<uint32_t>(hv[5]) << 16); hq.w = hv[6] | (static_cast<uint32_t>(hv[7]) << 16);
      *reinterpret_cast<uint4*>(hi + o) = hq;
      if (lo) {
        uint4 lq;
        lq.x = lv[0] | (static_cast<uint32_t>(lv[1]) << 16); lq.y = lv[2] | (static_cast<uint32_t>(lv[3]) << 16);
        lq.z = lv[4] | (static_cast<uint32_t>(lv[5]) << 16); lq.w = lv[6] | (static_cast<uint32_t>(lv[7]) << 16);
        *reinterpret_cast<uint4*>(lo + o) = lq;
      }
    }
  }
}

// ------------------------------------------------------------------------- fc / fc_seg + planes
// w [P*C, K] with source row j = c*P + pos (the reference's view(-1, C, 4, 4), model.py:379) is
// re-ordered to NHWC rows j' = pos*Cp + c.  A block owns (pos, 32 channels, 32 k): the tile is
// staged in shared memory and written as rows of the forward planes [P*Cp, Kp] (contiguous in k)
// and as columns of the transposed backward plane [Kp, P*Cp] (contiguous in c).
struct FcPlanes {
  void* fwd16;     // fp16 [P*Cp, Kp]          (no-grad forward), may be null
  void* fwd_hi;    // bf16 [P*Cp, Kp]          (hi/lo split forward), may be null
  void* fwd_lo;
  void* bwd;       // bf16 [Kp, P*Cp]          (data-gradient GEMM), may be null
};

__global__ void __launch_bounds__(256)
adam_pack_fc_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                    float* __restrict__ v, int C, int K, int P, int Cp, int Kp, const Hyper h,
                    const FcPlanes F) {
  __shared__ float s[32][33];
  const int k0 = blockIdx.x * 32, c0 = blockIdx.y * 32, pos = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float step_size = 0.f, isb2 = 0.f;
  if (g) {
    step_size = __ldg(h.lr) * __ldg(h.bc);
    isb2 = __ldg(h.bc + 1);
  }
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int cl = ty + rr * 8;
    const int c = c0 + cl, k = k0 + tx;
    float val = 0.f;
    if (c < C && k < K) {
      const int64_t o = (static_cast<int64_t>(c) * P + pos) * K + k;
      val = p[o];
      if (g) {
        float mm = m[o], vv = v[o];
        val = adam_update(val, g[o], mm, vv, step_size, isb2, h);
        p[o] = val;
        m[o] = mm;
        v[o] = vv;
      }
    }
    s[cl][tx] = val;
  }
  __syncthreads();
  // forward planes: row (pos*Cp + c), column k
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int cl = ty + rr * 8;
    const int c = c0 + cl, k = k0 + tx;
    if (c < Cp && k < Kp) {
      const float val = s[cl][tx];
      const int64_t o = (static_cast<int64_t>(pos) * Cp + c) * Kp + k;
      if (F.fwd16) static_cast<uint16_t*>(F.fwd16)[o] = __half_as_ushort(__float2half_rn(val));
      if (F.fwd_hi) {
        uint16_t hv, lv;
        split16(val, 1, hv, lv);
        static_cast<uint16_t*>(F.fwd_hi)[o] = hv;
        if (F.fwd_lo) static_cast<uint16_t*>(F.fwd_lo)[o] = lv;
      }
    }
  }
  // backward plane: row k, column (pos*Cp + c)
  if (F.bwd) {
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      const int kl = ty + rr * 8;
      const int k = k0 + kl, c = c0 + tx;
      if (c < Cp && k < Kp)
        static_cast<uint16_t*>(F.bwd)[static_cast<int64_t>(k) * (static_cast<int64_t>(P) * Cp) +
                                      static_cast<int64_t>(pos) * Cp + c] =
            __bfloat16_as_ushort(__float2bfloat16_rn(s[tx][kl]));
    }
  }
}

inline Hyper make_hyper(const cpcsv_adam_t* a) {
  Hyper h;
  h.lr = a->lr;
  h.bc = a->bc;
  h.beta1 = static_cast<float>(a->beta1);
  h.beta2 = static_cast<float>(a->beta2);
  h.omb1 = static_cast<float>(1.0 - a->beta1);
  h.omb2 = static_cast<float>(1.0 - a->beta2);
  h.eps = static_cast<float>(a->eps);
  return h;
}

}  // namespace
}  // namespace cpcsv

using namespace cpcsv;
#define STREAM(s) static_cast<cudaStream_t>(s)

extern "C" int cpcsv_adam_tick(float* step, double beta1, double beta2, float* bc, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(step && bc, "adam_tick: args");
  adam_tick_kernel<<<1, 1, 0, STREAM(stream)>>>(step, beta1, beta2, bc);
  return launched("adam_tick");
}

extern "C" int cpcsv_adam_multi(const cpcsv_adam_tensor_t* tensors, int32_t count, const cpcsv_adam_t* hyper,
                                cpcsv_stream_t stream) {
  CPCSV_REQUIRE(tensors && hyper && hyper->lr && hyper->bc && count > 0, "adam_multi: args");
  const Hyper h = make_hyper(hyper);
  for (int first = 0; first < count; first += kMaxTensors) {
    TensorTable T;
    const int n = count - first < kMaxTensors ? count - first : kMaxTensors;
    int64_t blocks = 0;
    for (int i = 0; i < n; ++i) {
      const cpcsv_adam_tensor_t& t = tensors[first + i];
      CPCSV_REQUIRE(t.p && t.g && t.m && t.v && t.n > 0, "adam_multi: tensor %d", first + i);
      T.p[i] = t.p; T.g[i] = t.g; T.m[i] = t.m; T.v[i] = t.v; T.n[i] = t.n;
      T.block0[i] = static_cast<int32_t>(blocks);
      blocks += ceil_div(t.n, kChunk);
      CPCSV_REQUIRE(blocks < (1ll << 31), "adam_multi: too many blocks");
    }
    T.block0[n] = static_cast<int32_t>(blocks);
    T.count = n;
    adam_multi_kernel<<<static_cast<unsigned>(blocks), 256, 0, STREAM(stream)>>>(T, h);
    int rc = launched("adam_multi");
    if (rc) return rc;
  }
  return 0;
}

extern "C" int cpcsv_adam_pack_conv(float* p, const float* g, float* m, float* v, int32_t Cout, int32_t Cin,
                                    int32_t kh, int32_t kw, const cpcsv_adam_t* hyper,
                                    const cpcsv_plane_t* planes, int32_t n_planes, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(p && Cout > 0 && Cin > 0 && (kh * kw == 9 || kh * kw == 16), "adam_pack_conv: 3x3 / 4x4 kernels only");
  CPCSV_REQUIRE(!g || (m && v && hyper && hyper->lr && hyper->bc), "adam_pack_conv: optimiser state missing");
  CPCSV_REQUIRE(n_planes >= 0 && n_planes <= CPCSV_MAX_PLANES && (n_planes == 0 || planes),
                "adam_pack_conv: %d planes", n_planes);
  CPCSV_REQUIRE(g || n_planes > 0, "adam_pack_conv: nothing to do");
  PlaneSet S;
  S.count = n_planes;
  int co_ext = Cout, ci_ext = Cin;
  for (int i = 0; i < n_planes; ++i) {
    const cpcsv_plane_t& P = planes[i];
    CPCSV_REQUIRE(P.hi && P.kind >= 0 && P.kind <= 3 && (P.dtype == 0 || P.dtype == 1), "adam_pack_conv: plane %d", i);
    CPCSV_REQUIRE(P.kind < 2 || (kh == 3 && kw == 3), "adam_pack_conv: sub-pixel merge needs 3x3");
    CPCSV_REQUIRE((P.kind >= 2) == (planes[0].kind >= 2), "adam_pack_conv: plain and merged planes of one weight");
    const bool tr = (P.kind == 1 || P.kind == 3);
    CPCSV_REQUIRE(P.rows_pad >= (tr ? Cin : Cout) && P.cols_pad >= (tr ? Cout : Cin) && P.cols_pad % 8 == 0,
                  "adam_pack_conv: plane %d padding", i);
    const int co_p = tr ? P.cols_pad : P.rows_pad, ci_p = tr ? P.rows_pad : P.cols_pad;
    if (co_p > co_ext) co_ext = co_p;
    if (ci_p > ci_ext) ci_ext = ci_p;
    S.pl[i] = P;
  }
  Hyper h = {};
  if (g) h = make_hyper(hyper);
  // tiles cover the PADDED index space so the zero padding of every plane is written too
  dim3 grid(static_cast<unsigned>(ceil_div(ci_ext, kTile)), static_cast<unsigned>(ceil_div(co_ext, kTile)));
  const bool merged = n_planes > 0 && planes[0].kind >= 2;
  if (merged) adam_pack_conv_kernel<9, true><<<grid, 256, 0, STREAM(stream)>>>(p, g, m, v, Cout, Cin, h, S);
  else if (kh * kw == 9) adam_pack_conv_kernel<9, false><<<grid, 256, 0, STREAM(stream)>>>(p, g, m, v, Cout, Cin, h, S);
  else adam_pack_conv_kernel<16, false><<<grid, 256, 0, STREAM(stream)>>>(p, g, m, v, Cout, Cin, h, S);
  return launched("adam_pack_conv");
}

extern "C" int cpcsv_adam_pack_fc(float* p, const float* g, float* m, float* v, int32_t C, int32_t K, int32_t P,
                                  int32_t Cp, int32_t Kp, const cpcsv_adam_t* hyper, void* fwd16, void* fwd_hi,
                                  void* fwd_lo, void* bwd, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(p && C > 0 && K > 0 && P > 0 && Cp >= C && Kp >= K, "adam_pack_fc: args");
  CPCSV_REQUIRE(!g || (m && v && hyper && hyper->lr && hyper->bc), "adam_pack_fc: optimiser state missing");
  CPCSV_REQUIRE(!fwd_lo || fwd_hi, "adam_pack_fc: lo without hi");
  Hyper h = {};
  if (g) h = make_hyper(hyper);
  FcPlanes F = {fwd16, fwd_hi, fwd_lo, bwd};
  dim3 grid(static_cast<unsigned>(ceil_div(Kp, 32)), static_cast<unsigned>(ceil_div(Cp, 32)),
            static_cast<unsigned>(P));
  adam_pack_fc_kernel<<<grid, 256, 0, STREAM(stream)>>>(p, g, m, v, C, K, P, Cp, Kp, h, F);
  return launched("adam_pack_fc");
}
