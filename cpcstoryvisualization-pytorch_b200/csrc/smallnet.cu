// fp32 kernels of the conditioning path (kept in fp32: SURVEY.md Appendix E item 3) and of the
// spectral-norm power iteration.  Problem sizes are tiny (B <= a few hundred rows, <= 1.8k
// columns): these are latency-bound, so each op is one compact launch.
#include "common.h"

namespace cpcsv {

// C[m][n] (+)= sum_k A(m,k) * B(n,k) + bias[n], with A(m,k) = A[m*sam + k*sak] and
// B(n,k) = B[n*sbn + k*sbk].  32x32 output tile, 256 threads (4 outputs each), K step 32.
// These GEMMs have a handful of output tiles and are latency-bound: the K range is split over
// gridDim.z (partial sums combined with atomicAdd into a pre-zeroed / accumulated C) so the
// whole chip works on one problem, and the next K tile is fetched into registers while the
// current one is multiplied.
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                     const float* __restrict__ B, int64_t sbn, int64_t sbk,
                     const float* __restrict__ bias, float* __restrict__ C, int64_t ldc, int M, int N,
                     int K, int accumulate, int k_chunk) {
  __shared__ float As[32][33];
  __shared__ float Bs[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // ty in 0..7
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);
  // thread -> tile element mapping that is contiguous in memory for either operand layout
  const bool a_kfast = (sak == 1), b_kfast = (sbk == 1);
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      const int a_m = a_kfast ? r : tx, a_k = a_kfast ? tx : r;
      const int m = m0 + a_m, ka = k0 + a_k;
      ra[i] = (m < M && ka < k_end) ? __ldg(A + m * sam + ka * sak) : 0.f;
      const int b_n = b_kfast ? r : tx, b_k = b_kfast ? tx : r;
      const int n = n0 + b_n, kb = k0 + b_k;
      rb[i] = (n < N && kb < k_end) ? __ldg(B + n * sbn + kb * sbk) : 0.f;
    }
  };
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (k_begin < k_end) fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i;
      As[a_kfast ? r : tx][a_kfast ? tx : r] = ra[i];
      Bs[b_kfast ? r : tx][b_kfast ? tx : r] = rb[i];
    }
    __syncthreads();
    if (k0 + 32 < k_end) fetch(k0 + 32);
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float b = Bs[tx][k];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(As[ty + 8 * i][k], b, acc[i]);
    }
    __syncthreads();
  }
  const int n = n0 + tx;
  if (n < N) {
    const float bv = (bias && blockIdx.z == 0) ? bias[n] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty + 8 * i;
      if (m < M) {
        const float v = acc[i] + bv;
        float* dst = C + static_cast<int64_t>(m) * ldc + n;
        if (gridDim.z > 1) {
          atomicAdd(dst, v);
        } else {
          *dst = accumulate ? (*dst + v) : v;
        }
      }
    }
  }
}

__global__ void zero_rows_kernel(float* __restrict__ C, int64_t ldc, int M, int N) {
  const int64_t total = static_cast<int64_t>(M) * N;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    C[(i / N) * ldc + (i % N)] = 0.f;
}

static int sgemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk,
                 const float* bias, float* C, int64_t ldc, int M, int N, int K, int accumulate,
                 cudaStream_t stream, const char* what) {
  CPCSV_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, "%s: args", what);
  dim3 grid(static_cast<unsigned>(ceil_div(N, 32)), static_cast<unsigned>(ceil_div(M, 32)));
  // split K until ~2 CTAs per SM are busy, keeping >= 2 K tiles per split
  const int64_t ctas = static_cast<int64_t>(grid.x) * grid.y;
  const int ksteps = static_cast<int>(ceil_div(K, 32));
  int splits = static_cast<int>(ceil_div(2 * num_sms(), ctas));
  if (splits > ksteps / 2) splits = ksteps / 2;
  if (splits < 1) splits = 1;
  const int k_chunk = static_cast<int>(ceil_div(ksteps, splits)) * 32;
  splits = static_cast<int>(ceil_div(K, k_chunk));
  grid.z = static_cast<unsigned>(splits);
  if (splits > 1 && !accumulate) {
    int64_t zb = ceil_div(static_cast<int64_t>(M) * N, 256);
    if (zb > 148) zb = 148;
    zero_rows_kernel<<<static_cast<unsigned>(zb), 256, 0, stream>>>(C, ldc, M, N);
  }
  sgemm_strided_kernel<<<grid, 256, 0, stream>>>(A, sam, sak, B, sbn, sbk, bias, C, ldc, M, N, K,
                                                 accumulate, k_chunk);
  return launched(what);
}

// Skinny case Y[M, N<=16] = X[M,K] W[N,K]^T with K in the thousands (D_GET_LOGITS' final 4x4
// conv and cate_classify are dot products over the 16k-element feature map): split K over
// the grid, reduce in the block, one atomicAdd per (row, output, split).  Y is pre-zeroed.
__global__ void __launch_bounds__(256)
linear_skinny_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W, int64_t ldw,
                     const float* __restrict__ bias, float* __restrict__ Y, int64_t ldy, int N, int K,
                     int kchunk) {
  const int m = blockIdx.y;
  const int k0 = blockIdx.x * kchunk;
  const int k1 = min(K, k0 + kchunk);
  float acc[16];
#pragma unroll
  for (int n = 0; n < 16; ++n) acc[n] = 0.f;
  const float* x = X + static_cast<int64_t>(m) * ldx;
  for (int k = k0 + threadIdx.x; k < k1; k += 256) {
    const float xv = x[k];
#pragma unroll
    for (int n = 0; n < 16; ++n)
      if (n < N) acc[n] = fmaf(xv, W[static_cast<int64_t>(n) * ldw + k], acc[n]);
  }
  __shared__ float sh[8][16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    float v = acc[n];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][n] = v;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += sh[w][threadIdx.x];
    if (bias && blockIdx.x == 0) v += bias[threadIdx.x];
    atomicAdd(Y + static_cast<int64_t>(m) * ldy + threadIdx.x, v);
  }
}

// ------------------------------------------------------------------------- GRU gates
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void gru_gates_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh,
                                     const float* __restrict__ h, int B, int H, float* __restrict__ hnew,
                                     float* __restrict__ save) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, j = i - b * H;
  const float* gib = gi + static_cast<int64_t>(b) * 3 * H;
  const float* ghb = gh + static_cast<int64_t>(b) * 3 * H;
  const float r = sigmoidf_(gib[j] + ghb[j]);
  const float z = sigmoidf_(gib[H + j] + ghb[H + j]);
  const float hn = ghb[2 * H + j];
  const float n = tanhf(gib[2 * H + j] + r * hn);
  hnew[i] = (1.f - z) * n + z * h[i];
  float* sv = save + static_cast<int64_t>(b) * 4 * H;
  sv[j] = r; sv[H + j] = z; sv[2 * H + j] = n; sv[3 * H + j] = hn;
}

__global__ void gru_gates_bwd_kernel(const float* __restrict__ dhnew, const float* __restrict__ h,
                                     const float* __restrict__ save, int B, int H,
                                     float* __restrict__ dgi, float* __restrict__ dgh,
                                     float* __restrict__ dh) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H) return;
  const int b = i / H, j = i - b * H;
  const float* sv = save + static_cast<int64_t>(b) * 4 * H;
  const float r = sv[j], z = sv[H + j], n = sv[2 * H + j], hn = sv[3 * H + j];
  const float d = dhnew[i];
  const float dn = d * (1.f - z);
  const float dz = d * (h[i] - n);
  const float dpre_n = dn * (1.f - n * n);
  const float dr = dpre_n * hn;
  const float dpre_r = dr * r * (1.f - r);
  const float dpre_z = dz * z * (1.f - z);
  float* dgib = dgi + static_cast<int64_t>(b) * 3 * H;
  float* dghb = dgh + static_cast<int64_t>(b) * 3 * H;
  dgib[j] = dpre_r; dgib[H + j] = dpre_z; dgib[2 * H + j] = dpre_n;
  dghb[j] = dpre_r; dghb[H + j] = dpre_z; dghb[2 * H + j] = dpre_n * r;
  dh[i] = d * z;
}

// ------------------------------------------------------------------------- CA_NET
__global__ void ca_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ eps, int B, int C,
                              float* __restrict__ mu, float* __restrict__ logvar,
                              float* __restrict__ code) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, j = i - b * C;
  const float m = fmaxf(pre[static_cast<int64_t>(b) * 2 * C + j], 0.f);
  const float lv = fmaxf(pre[static_cast<int64_t>(b) * 2 * C + C + j], 0.f);
  mu[i] = m;
  logvar[i] = lv;
  code[i] = eps[i] * expf(0.5f * lv) + m;
}

__global__ void ca_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ eps,
                              const float* __restrict__ dmu, const float* __restrict__ dlogvar,
                              const float* __restrict__ dcode, int B, int C, float* __restrict__ dpre) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, j = i - b * C;
  const float pm = pre[static_cast<int64_t>(b) * 2 * C + j];
  const float pl = pre[static_cast<int64_t>(b) * 2 * C + C + j];
  const float lv = fmaxf(pl, 0.f);
  const float dc = dcode ? dcode[i] : 0.f;
  const float gm = (dmu ? dmu[i] : 0.f) + dc;
  const float gl = (dlogvar ? dlogvar[i] : 0.f) + dc * eps[i] * 0.5f * expf(0.5f * lv);
  dpre[static_cast<int64_t>(b) * 2 * C + j] = pm > 0.f ? gm : 0.f;
  dpre[static_cast<int64_t>(b) * 2 * C + C + j] = pl > 0.f ? gl : 0.f;
}

// ------------------------------------------------------------------------- dynamic filter
__global__ void dfn1d_fwd_kernel(const float* __restrict__ img, const float* __restrict__ filt, int N,
                                 int C, int L, int K, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * L) return;
  const int n = i / L, x = i - n * L;
  const int pad = K / 2;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) {
    const float* im = img + (static_cast<int64_t>(n) * C + c) * L;
    const float* f = filt + (static_cast<int64_t>(n) * C + c) * K;
    for (int k = 0; k < K; ++k) {
      const int xi = x + k - pad;
      if (xi >= 0 && xi < L) acc = fmaf(im[xi], f[k], acc);
    }
  }
  out[i] = acc;
}

// one thread per dimg element and one per dfilt element (two index ranges in one launch)
__global__ void dfn1d_bwd_kernel(const float* __restrict__ img, const float* __restrict__ filt,
                                 const float* __restrict__ dout, int N, int C, int L, int K,
                                 float* __restrict__ dimg, float* __restrict__ dfilt) {
  const int64_t n_img = static_cast<int64_t>(N) * C * L;
  const int64_t n_flt = static_cast<int64_t>(N) * C * K;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int pad = K / 2;
  if (i < n_img) {
    const int xi = static_cast<int>(i % L);
    const int64_t nc = i / L;
    const int n = static_cast<int>(nc / C);
    const float* f = filt + nc * K;
    const float* d = dout + static_cast<int64_t>(n) * L;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const int x = xi - k + pad;
      if (x >= 0 && x < L) acc = fmaf(d[x], f[k], acc);
    }
    dimg[i] = acc;
  } else if (i < n_img + n_flt) {
    const int64_t j = i - n_img;
    const int k = static_cast<int>(j % K);
    const int64_t nc = j / K;
    const int n = static_cast<int>(nc / C);
    const float* im = img + nc * L;
    const float* d = dout + static_cast<int64_t>(n) * L;
    float acc = 0.f;
    for (int x = 0; x < L; ++x) {
      const int xi = x + k - pad;
      if (xi >= 0 && xi < L) acc = fmaf(d[x], im[xi], acc);
    }
    dfilt[j] = acc;
  }
}

__global__ void tanh_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = tanhf(x[i]);
}
__global__ void tanh_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                float* __restrict__ dx, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = dy[i] * (1.f - y[i] * y[i]);
}

// out = sigmoid(t * alpha + bias)   (final D_GET_LOGITS layer, reference model.py:79-80)
__global__ void affine_sigmoid_fwd_kernel(const float* __restrict__ t, const float* __restrict__ alpha,
                                          const float* __restrict__ bias, float* __restrict__ out,
                                          int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = alpha ? *alpha : 1.f, b = bias ? *bias : 0.f;
  out[i] = 1.f / (1.f + expf(-(t[i] * a + b)));
}
// dz = dout * out * (1 - out); dt = dz * alpha
__global__ void affine_sigmoid_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                          const float* __restrict__ alpha, float* __restrict__ dt,
                                          float* __restrict__ dz, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = alpha ? *alpha : 1.f;
  const float o = out[i];
  const float g = dout[i] * o * (1.f - o);
  dz[i] = g;
  dt[i] = g * a;
}

// ------------------------------------------------------------------------- spectral norm
// t[c] += sum_{r in chunk} W[r][c] * u[r]   (coalesced over c)
__global__ void sn_wt_u_kernel(const float* __restrict__ W, int R, int C, const float* __restrict__ u,
                               float* __restrict__ t) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int r0 = blockIdx.y * 32;
  const int r1 = min(R, r0 + 32);
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc = fmaf(W[static_cast<int64_t>(r) * C + c], u[r], acc);
  atomicAdd(&t[c], acc);
}
// t[r] = dot(W[r], v): one 256-thread block per row, 4 independent loads in flight per thread
__global__ void __launch_bounds__(256)
sn_w_v_kernel(const float* __restrict__ W, int R, int C, const float* __restrict__ v,
              float* __restrict__ t) {
  const int r = blockIdx.x;
  const float* w = W + static_cast<int64_t>(r) * C;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int c = threadIdx.x;
  for (; c + 768 < C; c += 1024) {
    a0 = fmaf(w[c], v[c], a0);
    a1 = fmaf(w[c + 256], v[c + 256], a1);
    a2 = fmaf(w[c + 512], v[c + 512], a2);
    a3 = fmaf(w[c + 768], v[c + 768], a3);
  }
  for (; c < C; c += 256) a0 = fmaf(w[c], v[c], a0);
  float acc = (a0 + a1) + (a2 + a3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += sh[i];
    t[r] = s;
  }
}
__device__ float block_sum(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = 0.f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t += sh[i];
  __syncthreads();
  return t;
}
// dst = src / max(||src||, eps) (single block)
__global__ void sn_normalize_kernel(const float* __restrict__ src, int n, float eps,
                                    float* __restrict__ dst) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fmaf(src[i], src[i], acc);
  const float nrm = sqrtf(block_sum(acc, sh));
  const float inv = 1.f / fmaxf(nrm, eps);
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i] * inv;
}
// t = W v is given.  power iteration: u = t / max(||t||, eps).  sigma = u . t
__global__ void sn_finish_kernel(const float* __restrict__ t, int R, float eps, int power, float* u,
                                 float* sigma, float* inv_sigma) {
  __shared__ float sh[32];
  if (power) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < R; i += blockDim.x) acc = fmaf(t[i], t[i], acc);
    const float nrm = sqrtf(block_sum(acc, sh));
    const float inv = 1.f / fmaxf(nrm, eps);
    for (int i = threadIdx.x; i < R; i += blockDim.x) u[i] = t[i] * inv;
    __syncthreads();
  }
  float acc = 0.f;
  for (int i = threadIdx.x; i < R; i += blockDim.x) acc = fmaf(u[i], t[i], acc);
  const float s = block_sum(acc, sh);
  if (threadIdx.x == 0) {
    *sigma = s;
    *inv_sigma = 1.f / s;
  }
}

// backward through W_eff = W / sigma, sigma = u^T W v:
//   dW = (G - (sum(G .* W) / sigma) * u v^T) / sigma
__global__ void sn_bwd_dot_kernel(const float* __restrict__ G, const float* __restrict__ W, int64_t n,
                                  float* __restrict__ acc_out) {
  __shared__ float sh[32];
  float acc = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    acc = fmaf(G[i], W[i], acc);
  const float s = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(acc_out, s);
}
// one block per row chunk (blockIdx.y walks the rows, blockIdx.x the columns): no per-element division;
// G and dW may be the same buffer
__global__ void __launch_bounds__(256)
sn_bwd_apply_kernel(const float* G, const float* __restrict__ u, const float* __restrict__ v,
                    const float* __restrict__ sigma, const float* __restrict__ gw, int R, int C, float* dW) {
  const float inv = 1.f / *sigma;
  const float coef = *gw * inv;
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const float vc = coef * v[c];
  for (int r = blockIdx.y; r < R; r += gridDim.y) {
    const int64_t i = static_cast<int64_t>(r) * C + c;
    dW[i] = (G[i] - u[r] * vc) * inv;
  }
}

// Zero fill as a kernel: inside the whole-step CUDA graph, memset NODES of independent branches
// were observed to execute in one shared order (a branch's first memset waited ~6 ms for memsets
// of two other branches), kernel nodes are not.
__global__ void zero_f32_kernel(float* __restrict__ p, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = 0.f;
}

static inline void zero_f32(float* p, int64_t n, cudaStream_t stream) {
  int64_t blocks = ceil_div(n, 256);
  if (blocks > 148) blocks = 148;
  zero_f32_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p, n);
}

}  // namespace cpcsv

using namespace cpcsv;
#define STREAM(s) static_cast<cudaStream_t>(s)

extern "C" int cpcsv_linear_f32(const float* X, int64_t ldx, const float* W, int64_t ldw,
                                const float* bias, float* Y, int64_t ldy, int32_t M, int32_t N,
                                int32_t K, int32_t accumulate, cpcsv_stream_t stream) {
  if (N <= 16 && K >= 2048 && ldy == N && X && W && Y && M > 0) {
    if (!accumulate) zero_f32(Y, static_cast<int64_t>(M) * N, STREAM(stream));
    const int kchunk = 2048;
    dim3 grid(static_cast<unsigned>(ceil_div(K, kchunk)), static_cast<unsigned>(M));
    linear_skinny_kernel<<<grid, 256, 0, STREAM(stream)>>>(X, ldx, W, ldw, bias, Y, ldy, N, K, kchunk);
    return launched("linear_f32/skinny");
  }
  return sgemm(X, ldx, 1, W, ldw, 1, bias, Y, ldy, M, N, K, accumulate, STREAM(stream), "linear_f32");
}
extern "C" int cpcsv_linear_tn_f32(const float* A, int64_t lda, const float* B, int64_t ldb, float* Y,
                                   int64_t ldy, int32_t M, int32_t N, int32_t K, int32_t accumulate,
                                   cpcsv_stream_t stream) {
  // Y[m][n] = sum_k A[k][m] * B[k][n]
  return sgemm(A, 1, lda, B, 1, ldb, nullptr, Y, ldy, M, N, K, accumulate, STREAM(stream),
               "linear_tn_f32");
}
extern "C" int cpcsv_linear_nn_f32(const float* X, int64_t ldx, const float* W, int64_t ldw, float* Y,
                                   int64_t ldy, int32_t M, int32_t N, int32_t K, int32_t accumulate,
                                   cpcsv_stream_t stream) {
  // Y[m][n] = sum_k X[m][k] * W[k][n]
  return sgemm(X, ldx, 1, W, 1, ldw, nullptr, Y, ldy, M, N, K, accumulate, STREAM(stream),
               "linear_nn_f32");
}

extern "C" int cpcsv_gru_gates_fwd(const float* gi, const float* gh, const float* h, int32_t B,
                                   int32_t H, float* hnew, float* save, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(gi && gh && h && hnew && save && B > 0 && H > 0, "gru_gates_fwd: args");
  gru_gates_fwd_kernel<<<static_cast<unsigned>(ceil_div(B * H, 256)), 256, 0, STREAM(stream)>>>(
      gi, gh, h, B, H, hnew, save);
  return launched("gru_gates_fwd");
}
extern "C" int cpcsv_gru_gates_bwd(const float* dhnew, const float* h, const float* save, int32_t B,
                                   int32_t H, float* dgi, float* dgh, float* dh,
                                   cpcsv_stream_t stream) {
  CPCSV_REQUIRE(dhnew && h && save && dgi && dgh && dh && B > 0 && H > 0, "gru_gates_bwd: args");
  gru_gates_bwd_kernel<<<static_cast<unsigned>(ceil_div(B * H, 256)), 256, 0, STREAM(stream)>>>(
      dhnew, h, save, B, H, dgi, dgh, dh);
  return launched("gru_gates_bwd");
}
extern "C" int cpcsv_ca_fwd(const float* pre, const float* eps, int32_t B, int32_t C, float* mu,
                            float* logvar, float* code, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(pre && eps && mu && logvar && code && B > 0 && C > 0, "ca_fwd: args");
  ca_fwd_kernel<<<static_cast<unsigned>(ceil_div(B * C, 256)), 256, 0, STREAM(stream)>>>(
      pre, eps, B, C, mu, logvar, code);
  return launched("ca_fwd");
}
extern "C" int cpcsv_ca_bwd(const float* pre, const float* eps, const float* dmu,
                            const float* dlogvar, const float* dcode, int32_t B, int32_t C,
                            float* dpre, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(pre && eps && dpre && B > 0 && C > 0, "ca_bwd: args");
  ca_bwd_kernel<<<static_cast<unsigned>(ceil_div(B * C, 256)), 256, 0, STREAM(stream)>>>(
      pre, eps, dmu, dlogvar, dcode, B, C, dpre);
  return launched("ca_bwd");
}
extern "C" int cpcsv_dfn1d_fwd(const float* img, const float* filt, int32_t N, int32_t C, int32_t L,
                               int32_t K, float* out, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(img && filt && out && N > 0 && C > 0 && L > 0 && K > 0, "dfn1d_fwd: args");
  dfn1d_fwd_kernel<<<static_cast<unsigned>(ceil_div(N * L, 128)), 128, 0, STREAM(stream)>>>(
      img, filt, N, C, L, K, out);
  return launched("dfn1d_fwd");
}
extern "C" int cpcsv_dfn1d_bwd(const float* img, const float* filt, const float* dout, int32_t N,
                               int32_t C, int32_t L, int32_t K, float* dimg, float* dfilt,
                               cpcsv_stream_t stream) {
  CPCSV_REQUIRE(img && filt && dout && dimg && dfilt && N > 0, "dfn1d_bwd: args");
  const int64_t work = static_cast<int64_t>(N) * C * (L + K);
  dfn1d_bwd_kernel<<<static_cast<unsigned>(ceil_div(work, 128)), 128, 0, STREAM(stream)>>>(
      img, filt, dout, N, C, L, K, dimg, dfilt);
  return launched("dfn1d_bwd");
}
extern "C" int cpcsv_tanh_fwd(const float* x, float* y, int64_t n, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(x && y && n > 0, "tanh_fwd: args");
  tanh_fwd_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, STREAM(stream)>>>(x, y, n);
  return launched("tanh_fwd");
}
extern "C" int cpcsv_tanh_bwd(const float* y, const float* dy, float* dx, int64_t n,
                              cpcsv_stream_t stream) {
  CPCSV_REQUIRE(y && dy && dx && n > 0, "tanh_bwd: args");
  tanh_bwd_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, STREAM(stream)>>>(y, dy, dx, n);
  return launched("tanh_bwd");
}

extern "C" int cpcsv_affine_sigmoid_fwd(const float* t, const float* alpha, const float* bias,
                                        float* out, int64_t n, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(t && out && n > 0, "affine_sigmoid_fwd: args");
  affine_sigmoid_fwd_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, STREAM(stream)>>>(
      t, alpha, bias, out, n);
  return launched("affine_sigmoid_fwd");
}
extern "C" int cpcsv_affine_sigmoid_bwd(const float* dout, const float* out, const float* alpha,
                                        float* dt, float* dz, int64_t n, cpcsv_stream_t stream) {
  CPCSV_REQUIRE(dout && out && dt && dz && n > 0, "affine_sigmoid_bwd: args");
  affine_sigmoid_bwd_kernel<<<static_cast<unsigned>(ceil_div(n, 256)), 256, 0, STREAM(stream)>>>(
      dout, out, alpha, dt, dz, n);
  return launched("affine_sigmoid_bwd");
}

extern "C" int cpcsv_spectral_sigma(const float* W, int32_t R, int32_t C, float* u, float* v,
                                    int32_t do_power_iteration, float eps, float* sigma,
                                    float* inv_sigma, float* scratch, cpcsv_stream_t stream_) {
  cudaStream_t stream = STREAM(stream_);
  CPCSV_REQUIRE(W && u && v && sigma && inv_sigma && scratch && R > 0 && C > 0, "spectral_sigma: args");
  float* tc = scratch;      // [C]
  float* tr = scratch + C;  // [R]
  if (do_power_iteration) {
    zero_f32(tc, C, stream);
    dim3 grid(static_cast<unsigned>(ceil_div(C, 128)), static_cast<unsigned>(ceil_div(R, 32)));
    sn_wt_u_kernel<<<grid, 128, 0, stream>>>(W, R, C, u, tc);
    int rc = launched("spectral_sigma/Wt_u");
    if (rc) return rc;
    sn_normalize_kernel<<<1, 1024, 0, stream>>>(tc, C, eps, v);
    rc = launched("spectral_sigma/normalize_v");
    if (rc) return rc;
  }
  sn_w_v_kernel<<<static_cast<unsigned>(R), 256, 0, stream>>>(W, R, C, v, tr);
  int rc = launched("spectral_sigma/W_v");
  if (rc) return rc;
  sn_finish_kernel<<<1, 1024, 0, stream>>>(tr, R, eps, do_power_iteration, u, sigma, inv_sigma);
  return launched("spectral_sigma/finish");
}

extern "C" int cpcsv_spectral_bwd(const float* G, const float* W, const float* u, const float* v,
                                  const float* sigma, int32_t R, int32_t C, float* dW,
                                  float* scratch, cpcsv_stream_t stream_) {
  cudaStream_t stream = STREAM(stream_);
  CPCSV_REQUIRE(G && W && u && v && sigma && dW && scratch && R > 0 && C > 0, "spectral_bwd: args");
  const int64_t n = static_cast<int64_t>(R) * C;
  zero_f32(scratch, 1, stream);
  int blocks = static_cast<int>(ceil_div(n, 256 * 8));
  if (blocks > num_sms() * 4) blocks = num_sms() * 4;
  if (blocks < 1) blocks = 1;
  sn_bwd_dot_kernel<<<blocks, 256, 0, stream>>>(G, W, n, scratch);
  int rc = launched("spectral_bwd/dot");
  if (rc) return rc;
  {
    const unsigned gx = static_cast<unsigned>(ceil_div(C, 256));
    int64_t gy = ceil_div(static_cast<int64_t>(num_sms()) * 8, gx);
    if (gy > R) gy = R;
    sn_bwd_apply_kernel<<<dim3(gx, static_cast<unsigned>(gy)), 256, 0, stream>>>(G, u, v, sigma, scratch, R, C, dW);
  }
  return launched("spectral_bwd/apply");
}

// second half alone: *dot = sum(G .* W) has already been accumulated (cpcsv_unpack_conv_wgrad_dot)
extern "C" int cpcsv_spectral_bwd_apply(const float* G, const float* u, const float* v, const float* sigma,
                                        const float* dot, int32_t R, int32_t C, float* dW,
                                        cpcsv_stream_t stream_) {
  cudaStream_t stream = STREAM(stream_);
  CPCSV_REQUIRE(G && u && v && sigma && dot && dW && R > 0 && C > 0, "spectral_bwd_apply: args");
  const unsigned gx = static_cast<unsigned>(ceil_div(C, 256));
  int64_t gy = ceil_div(static_cast<int64_t>(num_sms()) * 8, gx);
  if (gy > R) gy = R;
  sn_bwd_apply_kernel<<<dim3(gx, static_cast<unsigned>(gy)), 256, 0, stream>>>(G, u, v, sigma, dot, R, C, dW);
  return launched("spectral_bwd_apply");
}
