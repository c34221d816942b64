// fp32 mode of the path (BASELINE.json: per-step latent max-rel error <= 1e-4): every tensor stays fp32 in HBM and every
// contraction accumulates fp32 products on the CUDA cores (FFMA) — the bf16 tensor-core kernels of gemm2.cu cannot meet
// that tolerance (one bf16 store is 2^-9).  This is the accuracy mode, not the throughput mode: one generic tiled
// implicit-GEMM kernel (plain rows, 3x3 convolution gather with stride / asymmetric padding / nearest-2x up-sampling,
// two-level batching for the attention products), GroupNorm with double-precision statistics, LayerNorm, row soft-max,
// GEGLU and the small layout helpers.  Same C-ABI conventions as the rest of the library.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace edtr {

constexpr int kFBM = 64, kFBN = 64, kFBK = 16;

struct F32Gemm {
  const float* A;
  const float* W;
  float* C;
  int M, N, K;
  long long lda, ldw, ldc;
  int w_kn;                       // 1: W is [K, N] row-major (row stride ldw) instead of [N, K]
  int nb2;                        // blockIdx.z = b1 * nb2 + b2
  long long a_s1, a_s2, w_s1, w_s2, c_s1, c_s2;
  int conv;                       // 1: A is NHWC [B, H, W, Cin] (pixel stride lda), K = 9 * Cin (tap-major), rows = (b, yo, xo)
  int H, W_, Cin, Ho, Wo, stride, pad_t, pad_l, up2x;
  float alpha;
  const float* bias;
  const float* rowvec;
  long long rowvec_ld;
  int rows_per_group;
  const float* residual;
  long long ldr;
  int act;                        // EDTR_ACT_NONE / EDTR_ACT_SILU
  int out_nchw;                   // 1: C[((row / hw) * N + col) * hw + row % hw]
  int hw;
};

__device__ __forceinline__ float f32_silu(float x) { return x / (1.f + __expf(-x)); }

// One element of the (implicit) A matrix.
__device__ __forceinline__ float f32_load_a(const F32Gemm& p, const float* A, int row, int k, int rb, int ry, int rx) {
  if (!p.conv) return A[static_cast<long long>(row) * p.lda + k];
  const int tap = k / p.Cin, c = k - tap * p.Cin;
  const int ty = tap / 3, tx = tap - ty * 3;
  int y, x;
  if (p.up2x) {   // 3x3 / pad 1 on the nearest-2x up-sampled grid: source pixel (yu >> 1, xu >> 1)
    const int yu = ry + ty - 1, xu = rx + tx - 1;
    if (yu < 0 || xu < 0 || yu >= 2 * p.H || xu >= 2 * p.W_) return 0.f;
    y = yu >> 1;
    x = xu >> 1;
  } else {
    y = ry * p.stride + ty - p.pad_t;
    x = rx * p.stride + tx - p.pad_l;
    if (y < 0 || x < 0 || y >= p.H || x >= p.W_) return 0.f;
  }
  return A[((static_cast<long long>(rb) * p.H + y) * p.W_ + x) * p.lda + c];
}

__global__ void __launch_bounds__(256)
f32_gemm_kernel(const F32Gemm p) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  __shared__ float As[kFBK][kFBM + 4];
  __shared__ float Ws[kFBK][kFBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kFBM, n0 = blockIdx.x * kFBN;
  const int b1 = blockIdx.z / p.nb2, b2 = blockIdx.z - b1 * p.nb2;
  const float* A = p.A + b1 * p.a_s1 + b2 * p.a_s2;
  const float* W = p.W + b1 * p.w_s1 + b2 * p.w_s2;
  float* C = p.C + b1 * p.c_s1 + b2 * p.c_s2;
  // loader mapping: A tile row = tid / 4, four consecutive k; W tile row (n) = tid / 4, four consecutive k
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int arow = m0 + lr;
  int rb = 0, ry = 0, rx = 0;
  if (p.conv && arow < p.M) {
    const int hw = p.Ho * p.Wo;
    rb = arow / hw;
    const int r = arow - rb * hw;
    ry = r / p.Wo;
    rx = r - ry * p.Wo;
  }
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += kFBK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + lk + e;
      As[lk + e][lr] = (arow < p.M && k < p.K) ? f32_load_a(p, A, arow, k, rb, ry, rx) : 0.f;
    }
    if (!p.w_kn) {
      const int n = n0 + lr;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + lk + e;
        Ws[lk + e][lr] = (n < p.N && k < p.K) ? W[static_cast<long long>(n) * p.ldw + k] : 0.f;
      }
    } else {
      const int kr = tid >> 4, nn = (tid & 15) * 4;   // 16 k rows x 64 n
      const int k = k0 + kr;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = n0 + nn + e;
        Ws[kr][nn + e] = (n < p.N && k < p.K) ? W[static_cast<long long>(k) * p.ldw + n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kFBK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (p.bias != nullptr) v += p.bias[col];
      if (p.rowvec != nullptr) v += p.rowvec[static_cast<long long>(row / p.rows_per_group) * p.rowvec_ld + col];
      if (p.residual != nullptr) v += p.residual[static_cast<long long>(row) * p.ldr + col];
      if (p.act == EDTR_ACT_SILU) v = f32_silu(v);
      if (p.out_nchw) {
        const int img = row / p.hw;
        C[(static_cast<long long>(img) * p.N + col) * p.hw + (row - img * p.hw)] = v;
      } else {
        C[static_cast<long long>(row) * p.ldc + col] = v;
      }
    }
  }
}


// Second form of the same contraction for the shapes that allow 16-byte operand loads (K % 4 == 0, strides % 4 == 0,
// 16-byte aligned bases; convolution: Cin % 4 == 0; W as [N, K]): 128 x 64 x 16 tiles, 8 x 4 outputs per thread, the
// next tile's global loads are issued into registers before the current tile's arithmetic (software pipelining), two
// float4 of A and one of W per thread and tile.  Same epilogue, same summation order inside a tile row.
constexpr int kGBM = 128;

__device__ __forceinline__ float4 f32_load_a4(const F32Gemm& p, const float* A, int row, int k, int rb, int ry, int rx) {
  // four consecutive k of one row; in convolution mode they lie inside one tap (Cin % 4 == 0, k % 4 == 0)
  if (!p.conv) return __ldg(reinterpret_cast<const float4*>(A + static_cast<long long>(row) * p.lda + k));
  const int tap = k / p.Cin, c = k - tap * p.Cin;
  const int ty = tap / 3, tx = tap - ty * 3;
  int y, x;
  if (p.up2x) {
    const int yu = ry + ty - 1, xu = rx + tx - 1;
    if (yu < 0 || xu < 0 || yu >= 2 * p.H || xu >= 2 * p.W_) return make_float4(0.f, 0.f, 0.f, 0.f);
    y = yu >> 1;
    x = xu >> 1;
  } else {
    y = ry * p.stride + ty - p.pad_t;
    x = rx * p.stride + tx - p.pad_l;
    if (y < 0 || x < 0 || y >= p.H || x >= p.W_) return make_float4(0.f, 0.f, 0.f, 0.f);
  }
  return __ldg(reinterpret_cast<const float4*>(A + ((static_cast<long long>(rb) * p.H + y) * p.W_ + x) * p.lda + c));
}

__global__ void __launch_bounds__(256)
f32_gemm_vec_kernel(const F32Gemm p) {
  PdlScope pdl_scope;
  __shared__ __align__(16) float As[kFBK][kGBM + 4];
  __shared__ __align__(16) float Ws[kFBK][kFBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kGBM, n0 = blockIdx.x * kFBN;
  const int b1 = blockIdx.z / p.nb2, b2 = blockIdx.z - b1 * p.nb2;
  const float* A = p.A + b1 * p.a_s1 + b2 * p.a_s2;
  const float* W = p.W + b1 * p.w_s1 + b2 * p.w_s2;
  float* C = p.C + b1 * p.c_s1 + b2 * p.c_s2;
  // loaders: A row = tid / 2, k offsets (tid & 1) * 8 + {0, 4}; W row (n) = tid / 4, k offset (tid & 3) * 4
  const int ar = tid >> 1, ak = (tid & 1) * 8;
  const int wr = tid >> 2, wk = (tid & 3) * 4;
  const int arow = m0 + ar;
  const bool a_ok = arow < p.M, w_ok = n0 + wr < p.N;
  int rb = 0, ry = 0, rx = 0;
  if (p.conv && a_ok) {
    const int hw = p.Ho * p.Wo;
    rb = arow / hw;
    const int r = arow - rb * hw;
    ry = r / p.Wo;
    rx = r - ry * p.Wo;
  }
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto fetch = [&](int k0, float4& a0, float4& a1, float4& w0) {
    a0 = (a_ok && k0 + ak < p.K) ? f32_load_a4(p, A, arow, k0 + ak, rb, ry, rx) : z4;
    a1 = (a_ok && k0 + ak + 4 < p.K) ? f32_load_a4(p, A, arow, k0 + ak + 4, rb, ry, rx) : z4;
    w0 = (w_ok && k0 + wk < p.K) ? __ldg(reinterpret_cast<const float4*>(W + static_cast<long long>(n0 + wr) * p.ldw + k0 + wk)) : z4;
  };
  auto stash = [&](const float4& a0, const float4& a1, const float4& w0) {
    As[ak + 0][ar] = a0.x; As[ak + 1][ar] = a0.y; As[ak + 2][ar] = a0.z; As[ak + 3][ar] = a0.w;
    As[ak + 4][ar] = a1.x; As[ak + 5][ar] = a1.y; As[ak + 6][ar] = a1.z; As[ak + 7][ar] = a1.w;
    Ws[wk + 0][wr] = w0.x; Ws[wk + 1][wr] = w0.y; Ws[wk + 2][wr] = w0.z; Ws[wk + 3][wr] = w0.w;
  };
  const int ty = tid >> 4, tx = tid & 15;      // outputs: rows ty * 8 .. + 7, columns tx * 4 .. + 3
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float4 a0, a1, w0;
  fetch(0, a0, a1, w0);
  stash(a0, a1, w0);
  __syncthreads();
  for (int k0 = 0; k0 < p.K; k0 += kFBK) {
    const bool more = k0 + kFBK < p.K;
    if (more) fetch(k0 + kFBK, a0, a1, w0);      // in flight during the arithmetic below
#pragma unroll
    for (int k = 0; k < kFBK; ++k) {
      const float4 x0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 x1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float av[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
    if (more) {
      stash(a0, a1, w0);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + ty * 8 + i;
    if (row >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (p.bias != nullptr) v += p.bias[col];
      if (p.rowvec != nullptr) v += p.rowvec[static_cast<long long>(row / p.rows_per_group) * p.rowvec_ld + col];
      if (p.residual != nullptr) v += p.residual[static_cast<long long>(row) * p.ldr + col];
      if (p.act == EDTR_ACT_SILU) v = f32_silu(v);
      if (p.out_nchw) {
        const int img = row / p.hw;
        C[(static_cast<long long>(img) * p.N + col) * p.hw + (row - img * p.hw)] = v;
      } else {
        C[static_cast<long long>(row) * p.ldc + col] = v;
      }
    }
  }
}

// ------------------------------------------------------------------ GroupNorm (fp32, double-precision statistics)
// pass 1: grid (chunks, B); thread <-> channel, per-channel (sum, sum of squares) over the chunk's rows in double.
__global__ void __launch_bounds__(256)
f32_gn_stats_kernel(const float* __restrict__ X, long long ldx, int HW, int C, int rows, double* __restrict__ part) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const int b = blockIdx.y, ch = blockIdx.x;
  const int p0 = ch * rows, p1 = min(HW, p0 + rows);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double s = 0.0, ss = 0.0;
    const float* x = X + (static_cast<long long>(b) * HW + p0) * ldx + c;
    for (int r = p0; r < p1; ++r, x += ldx) {
      const double v = static_cast<double>(*x);
      s += v;
      ss += v * v;
    }
    double* dst = part + ((static_cast<long long>(b) * gridDim.x + ch) * C + c) * 2;
    dst[0] = s;
    dst[1] = ss;
  }
}
// pass 2: one thread per (image, group): mean / rstd from the per-channel partial sums (fixed order)
__global__ void f32_gn_finalize_kernel(const double* __restrict__ part, int B, int chunks, int C, int groups, int HW,
                                       float eps, float* __restrict__ mean_rstd) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * groups) return;
  const int b = i / groups, g = i - b * groups, cpg = C / groups;
  double s = 0.0, ss = 0.0;
  for (int ch = 0; ch < chunks; ++ch)
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const double* src = part + ((static_cast<long long>(b) * chunks + ch) * C + c) * 2;
      s += src[0];
      ss += src[1];
    }
  const double n = static_cast<double>(cpg) * HW;
  const double mean = s / n;
  double var = ss / n - mean * mean;
  if (var < 0.0) var = 0.0;
  mean_rstd[2 * i] = static_cast<float>(mean);
  mean_rstd[2 * i + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}
// pass 3: elementwise normalise (+ SiLU)
__global__ void __launch_bounds__(256)
f32_gn_apply_kernel(const float* __restrict__ X, long long ldx, float* __restrict__ Y, long long ldy, int HW, int C,
                    int groups, const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, int silu, long long total) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  const long long row = idx / C;
  const int b = static_cast<int>(row / HW);
  const int g = c / (C / groups);
  const float mean = mean_rstd[2 * (b * groups + g)], rstd = mean_rstd[2 * (b * groups + g) + 1];
  float v = (X[row * ldx + c] - mean) * rstd * gamma[c] + beta[c];
  if (silu) v = f32_silu(v);
  Y[row * ldy + c] = v;
}

// ------------------------------------------------------------------ LayerNorm (fp32): one warp per row, two exact passes
__global__ void __launch_bounds__(256)
f32_layernorm_kernel(const float* __restrict__ X, long long ldx, float* __restrict__ Y, long long ldy, int M, int C,
                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* x = X + static_cast<long long>(row) * ldx;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += x[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / static_cast<float>(C);
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = x[c] - mean;
    ss = fmaf(d, d, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / static_cast<float>(C) + eps);
  float* y = Y + static_cast<long long>(row) * ldy;
  for (int c = lane; c < C; c += 32) y[c] = (x[c] - mean) * rstd * gamma[c] + beta[c];
}

// ------------------------------------------------------------------ row soft-max in place (fp32): one CTA per row
__global__ void __launch_bounds__(256)
f32_softmax_rows_kernel(float* __restrict__ S, long long lds, int N, float scale) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  __shared__ float red[32];
  float* s = S + static_cast<long long>(blockIdx.x) * lds;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < N; i += blockDim.x) mx = fmaxf(mx, s[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = lane < nw ? red[lane] : -INFINITY;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float e = expf((s[i] - mx) * scale);
    s[i] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = lane < nw ? red[lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  for (int i = threadIdx.x; i < N; i += blockDim.x) s[i] *= inv;
}

// ------------------------------------------------------------------ elementwise helpers
// GEGLU: Y[m, n] = X[m, n] * gelu_erf(X[m, N + n])  (model/attention.py:20-27: chunk order [x | gate])
__global__ void f32_geglu_kernel(const float* __restrict__ X, long long ldx, float* __restrict__ Y, long long ldy, int N,
                                 long long total) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long m = idx / N;
  const int n = static_cast<int>(idx - m * N);
  const float x = X[m * ldx + n], g = X[m * ldx + N + n];
  Y[m * ldy + n] = x * (0.5f * g * (1.f + erff(g * 0.70710678118654752f)));
}
__global__ void f32_silu_kernel(const float* __restrict__ X, float* __restrict__ Y, long long total) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx < total) Y[idx] = f32_silu(X[idx]);
}
// [B, C, HW] fp32 -> channels [coff, coff + C) of [B, HW, ldy] fp32
__global__ void f32_nchw_to_nhwc_kernel(const float* __restrict__ X, float* __restrict__ Y, long long ldy, int C, int HW,
                                        int coff, float scale, long long total) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  const long long row = idx / C;
  const long long b = row / HW, pix = row - b * HW;
  Y[row * ldy + coff + c] = X[(b * C + c) * HW + pix] * scale;
}
// timestep embedding [cos | sin] in fp32 (model/util.py:98-118)
__global__ void f32_timestep_embedding_kernel(const long long* __restrict__ t, float* __restrict__ out, int dim,
                                              float max_period, int B) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (i >= B * half) return;
  const int b = i / half, k = i - b * half;
  const float f = expf(-logf(max_period) * static_cast<float>(k) / static_cast<float>(half));
  const float a = static_cast<float>(t[b]) * f;
  out[static_cast<long long>(b) * dim + k] = cosf(a);
  out[static_cast<long long>(b) * dim + half + k] = sinf(a);
}

}  // namespace edtr

using namespace edtr;

extern "C" int edtr_f32_gemm(const EdtrF32Gemm* g, void* stream) {
  EDTR_REQUIRE(g != nullptr && g->A && g->W && g->C, "A/W/C is NULL");
  EDTR_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0 && g->batch1 > 0 && g->batch2 > 0, "bad fp32 GEMM shape");
  EDTR_REQUIRE(static_cast<long long>(g->batch1) * g->batch2 <= 65535, "too many batches");
  EDTR_REQUIRE(g->act == EDTR_ACT_NONE || g->act == EDTR_ACT_SILU, "fp32 GEMM epilogue: act must be NONE or SILU");
  F32Gemm p{};
  p.A = g->A; p.W = g->W; p.C = g->C;
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.lda = g->lda; p.ldw = g->ldw; p.ldc = g->ldc;
  p.w_kn = g->w_kn;
  p.nb2 = g->batch2;
  p.a_s1 = g->a_stride1; p.a_s2 = g->a_stride2; p.w_s1 = g->w_stride1; p.w_s2 = g->w_stride2;
  p.c_s1 = g->c_stride1; p.c_s2 = g->c_stride2;
  p.conv = g->conv;
  if (g->conv) {
    EDTR_REQUIRE(g->H > 0 && g->W_in > 0 && g->Cin > 0 && g->Ho > 0 && g->Wo > 0 && g->K == 9 * g->Cin,
                 "fp32 convolution: K must be 9 * Cin and the geometry positive");
    EDTR_REQUIRE(g->M % (g->Ho * g->Wo) == 0 && (g->up2x || g->conv_stride >= 1), "fp32 convolution: bad row count / stride");
    EDTR_REQUIRE(!g->up2x || (g->Ho == 2 * g->H && g->Wo == 2 * g->W_in), "up2x output grid must be 2H x 2W");
    p.H = g->H; p.W_ = g->W_in; p.Cin = g->Cin; p.Ho = g->Ho; p.Wo = g->Wo; p.stride = g->conv_stride;
    p.pad_t = g->pad_top; p.pad_l = g->pad_left; p.up2x = g->up2x;
  }
  p.alpha = g->alpha; p.bias = g->bias; p.rowvec = g->rowvec; p.rowvec_ld = g->rowvec_ld;
  p.rows_per_group = g->rows_per_group > 0 ? g->rows_per_group : 1;
  p.residual = g->residual; p.ldr = g->ldr; p.act = g->act; p.out_nchw = g->out_nchw; p.hw = g->hw;
  EDTR_REQUIRE(!g->out_nchw || (g->hw > 0 && g->M % g->hw == 0), "NCHW output needs hw | M");
  // shapes that allow 16-byte operand loads run the 128 x 64 software-pipelined form (measured on B200: VAE decode
  // 20.8 -> 31.8 TFLOP/s, fp32 restore 2.43 -> 3.16 img/s at B = 4); EDTR_F32_GEMM_VEC=0 keeps the generic kernel (A/B)
  static const bool vec_enabled = [] {
    const char* e = getenv("EDTR_F32_GEMM_VEC");
    return e == nullptr || e[0] != '0';
  }();
  const auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool strides_ok = g->a_stride1 % 4 == 0 && g->a_stride2 % 4 == 0 && g->w_stride1 % 4 == 0 && g->w_stride2 % 4 == 0;
  const bool vec = vec_enabled && !g->w_kn && g->K % 4 == 0 && g->lda % 4 == 0 && g->ldw % 4 == 0 && al16(g->A) && al16(g->W) &&
                   strides_ok && (!g->conv || g->Cin % 4 == 0) && g->M >= kGBM;
  if (vec) {
    dim3 grid((g->N + kFBN - 1) / kFBN, (g->M + kGBM - 1) / kGBM, g->batch1 * g->batch2);
    EDTR_REQUIRE(grid.y <= 65535, "M too large for the fp32 GEMM grid");
    EDTR_LAUNCH(f32_gemm_vec_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), p);
    return check_launch("f32_gemm_vec_kernel");
  }
  dim3 grid((g->N + kFBN - 1) / kFBN, (g->M + kFBM - 1) / kFBM, g->batch1 * g->batch2);
  EDTR_REQUIRE(grid.y <= 65535, "M too large for the fp32 GEMM grid");
  EDTR_LAUNCH(f32_gemm_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), p);
  return check_launch("f32_gemm_kernel");
}

extern "C" size_t edtr_f32_groupnorm_scratch_bytes(int B, int HW, int C) {
  if (B <= 0 || HW <= 0 || C <= 0) return 0;
  const int chunks = HW >= 4096 ? 64 : (HW >= 64 ? HW / 64 : 1);
  return static_cast<size_t>(B) * chunks * C * 2 * sizeof(double) + static_cast<size_t>(B) * 64 * 2 * sizeof(float);
}

extern "C" int edtr_f32_groupnorm(const float* X, long long ldx, float* Y, long long ldy, int B, int HW, int C, int groups,
                                  const float* gamma, const float* beta, float eps, int silu, void* scratch,
                                  void* stream) {
  EDTR_REQUIRE(X && Y && gamma && beta && scratch, "X/Y/gamma/beta/scratch is NULL");
  EDTR_REQUIRE(B > 0 && HW > 0 && C > 0 && groups > 0 && groups <= 64 && C % groups == 0 && B <= 65535, "bad GroupNorm shape");
  EDTR_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 7) == 0, "scratch must be 8-byte aligned");
  const int chunks = HW >= 4096 ? 64 : (HW >= 64 ? HW / 64 : 1);
  const int rows = (HW + chunks - 1) / chunks;
  double* part = static_cast<double*>(scratch);
  float* mr = reinterpret_cast<float*>(part + static_cast<size_t>(B) * chunks * C * 2);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  EDTR_LAUNCH(f32_gn_stats_kernel, dim3(chunks, B), 256, 0, st, X, ldx, HW, C, rows, part);
  EDTR_LAUNCH(f32_gn_finalize_kernel, (B * groups + 127) / 128, 128, 0, st, part, B, chunks, C, groups, HW, eps, mr);
  const long long total = static_cast<long long>(B) * HW * C;
  EDTR_LAUNCH(f32_gn_apply_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, st, X, ldx, Y, ldy, HW, C, groups, mr,
              gamma, beta, silu, total);
  return check_launch("f32_groupnorm");
}

extern "C" int edtr_f32_layernorm(const float* X, long long ldx, float* Y, long long ldy, int M, int C, const float* gamma,
                                  const float* beta, float eps, void* stream) {
  EDTR_REQUIRE(X && Y && gamma && beta && M > 0 && C > 0, "bad LayerNorm arguments");
  EDTR_LAUNCH(f32_layernorm_kernel, (M + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream), X, ldx, Y, ldy, M, C, gamma,
              beta, eps);
  return check_launch("f32_layernorm_kernel");
}

extern "C" int edtr_f32_softmax_rows(float* S, long long lds, long long rows, int N, float scale, void* stream) {
  EDTR_REQUIRE(S && rows > 0 && rows < (1ll << 31) && N > 0 && lds >= N, "bad soft-max arguments");
  EDTR_LAUNCH(f32_softmax_rows_kernel, static_cast<unsigned>(rows), 256, 0, static_cast<cudaStream_t>(stream), S, lds, N,
              scale);
  return check_launch("f32_softmax_rows_kernel");
}

extern "C" int edtr_f32_geglu(const float* X, long long ldx, float* Y, long long ldy, long long M, int N, void* stream) {
  EDTR_REQUIRE(X && Y && M > 0 && N > 0 && ldx >= 2 * N && ldy >= N, "bad GEGLU arguments");
  const long long total = M * N;
  EDTR_LAUNCH(f32_geglu_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream), X, ldx, Y,
              ldy, N, total);
  return check_launch("f32_geglu_kernel");
}

extern "C" int edtr_f32_silu(const float* X, float* Y, long long n, void* stream) {
  EDTR_REQUIRE(X && Y && n > 0, "bad SiLU arguments");
  EDTR_LAUNCH(f32_silu_kernel, static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream), X, Y, n);
  return check_launch("f32_silu_kernel");
}

extern "C" int edtr_f32_nchw_to_nhwc(const float* X, float* Y, long long ldy, int B, int C, int HW, int coff, float scale,
                                     void* stream) {
  EDTR_REQUIRE(X && Y && B > 0 && C > 0 && HW > 0 && coff >= 0 && ldy >= coff + C, "bad layout arguments");
  const long long total = static_cast<long long>(B) * HW * C;
  EDTR_LAUNCH(f32_nchw_to_nhwc_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream), X,
              Y, ldy, C, HW, coff, scale, total);
  return check_launch("f32_nchw_to_nhwc_kernel");
}

extern "C" int edtr_f32_timestep_embedding(const long long* t, float* out, int B, int dim, float max_period, void* stream) {
  EDTR_REQUIRE(t && out && B > 0 && dim > 0 && dim % 2 == 0, "bad timestep embedding arguments");
  EDTR_LAUNCH(f32_timestep_embedding_kernel, (B * dim / 2 + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), t, out, dim,
              max_period, B);
  return check_launch("f32_timestep_embedding_kernel");
}
