// Layout / gather / pointwise kernels around the tensor-core path: NCHW<->channels-
// last conversion, nearest x2 upsample, im2col for strided convolutions, timestep
// embedding, and the fused spaced-DDPM sampler update.
#include "common.cuh"
#include "ptx.cuh"

namespace edtr {

// Y[(b,p), coff+c] = X[b,c,p]; one thread per pixel, coalesced reads per channel plane.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ X, __nv_bfloat16* __restrict__ Y, int ldy,
                                    int coff, int C, int HW, size_t total_pix) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total_pix) return;
  const size_t b = i / HW, pix = i - b * HW;
  const float* x = X + b * C * HW + pix;
  __nv_bfloat16* y = Y + i * ldy + coff;
  for (int c = 0; c < C; ++c) y[c] = __float2bfloat16(__ldg(x + static_cast<size_t>(c) * HW));
}

// Y[(b,p), coff+co] = bias[co] + sum_ci W[co,ci] * (scale * X[b,ci,p]); tiny channel counts, fp32 math.
__global__ void pointwise_nchw_kernel(const float* __restrict__ X, const float* __restrict__ Wm,
                                      const float* __restrict__ bias, float scale,
                                      __nv_bfloat16* __restrict__ Y, int ldy, int coff, int Cin, int Cout, int HW,
                                      size_t total_pix) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total_pix) return;
  const size_t b = i / HW, pix = i - b * HW;
  const float* x = X + b * Cin * HW + pix;
  float xin[16];
  for (int c = 0; c < Cin; ++c) xin[c] = scale * __ldg(x + static_cast<size_t>(c) * HW);
  __nv_bfloat16* y = Y + i * ldy + coff;
  for (int co = 0; co < Cout; ++co) {
    float acc = bias != nullptr ? __ldg(bias + co) : 0.f;
    for (int c = 0; c < Cin; ++c) acc += __ldg(Wm + co * Cin + c) * xin[c];
    y[co] = __float2bfloat16(acc);
  }
}

// Tiled transpose [B, HW, C] (bf16, stride ldx) -> [B, C, HW] (fp32 or bf16).
template <typename OutT>
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ X, int ldx, OutT* __restrict__ Y, int C,
                                    int HW) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int pix = p0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (pix < HW && c < C)
                               ? __bfloat162float(X[(static_cast<size_t>(b) * HW + pix) * ldx + c])
                               : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, pix = p0 + threadIdx.x;
    if (c < C && pix < HW) {
      const float v = tile[threadIdx.x][r];
      if constexpr (sizeof(OutT) == 4) Y[(static_cast<size_t>(b) * C + c) * HW + pix] = v;
      else Y[(static_cast<size_t>(b) * C + c) * HW + pix] = __float2bfloat16(v);
    }
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ X, __nv_bfloat16* __restrict__ Y, size_t n) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) Y[i] = __float2bfloat16(X[i]);
}

// One thread per (output pixel, 16-byte channel vector).
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ X, int ldx, __nv_bfloat16* __restrict__ Y,
                                  int ldy, int H, int W, int vpr, size_t total) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = static_cast<int>(i % vpr);
  const size_t opix = i / vpr;
  const int W2 = 2 * W, H2 = 2 * H;
  const int ox = static_cast<int>(opix % W2);
  const int oy = static_cast<int>((opix / W2) % H2);
  const size_t b = opix / (static_cast<size_t>(W2) * H2);
  const size_t ipix = (b * H + (oy >> 1)) * W + (ox >> 1);
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(X + ipix * ldx + v * 8));
  *reinterpret_cast<uint4*>(Y + opix * ldy + v * 8) = u;
}

// One thread per (output row, tap, 16-byte channel vector).
__global__ void im2col_kernel(const __nv_bfloat16* __restrict__ X, int ldx, __nv_bfloat16* __restrict__ Y,
                              int H, int W, int vpr, int KH, int KW, int stride, int pad_top, int pad_left,
                              int Ho, int Wo, size_t total) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int v = static_cast<int>(i % vpr);
  size_t t = i / vpr;
  const int taps = KH * KW;
  const int tap = static_cast<int>(t % taps);
  const size_t orow = t / taps;
  const int ox = static_cast<int>(orow % Wo);
  const int oy = static_cast<int>((orow / Wo) % Ho);
  const size_t b = orow / (static_cast<size_t>(Wo) * Ho);
  const int iy = oy * stride + tap / KW - pad_top;
  const int ix = ox * stride + tap % KW - pad_left;
  uint4 u = make_uint4(0, 0, 0, 0);
  if (iy >= 0 && iy < H && ix >= 0 && ix < W)
    u = __ldg(reinterpret_cast<const uint4*>(X + ((b * H + iy) * W + ix) * ldx + v * 8));
  *reinterpret_cast<uint4*>(Y + (orow * taps + tap) * (static_cast<size_t>(vpr) * 8) + v * 8) = u;
}

__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, __nv_bfloat16* __restrict__ Y, int dim,
                                          float log_max_period) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const int b = blockIdx.x;
  const int half = dim / 2;
  const float tv = static_cast<float>(t[b]);
  for (int k = threadIdx.x; k < half; k += blockDim.x) {
    const float freq = expf(-log_max_period * static_cast<float>(k) / static_cast<float>(half));
    const float a = tv * freq;
    Y[static_cast<size_t>(b) * dim + k] = __float2bfloat16(cosf(a));
    Y[static_cast<size_t>(b) * dim + half + k] = __float2bfloat16(sinf(a));
  }
  if ((dim & 1) && threadIdx.x == 0) Y[static_cast<size_t>(b) * dim + dim - 1] = __float2bfloat16(0.f);
}

__global__ void sampler_update_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                                      const float* __restrict__ noise, const int64_t* __restrict__ index,
                                      const float* __restrict__ sqrt_recip, const float* __restrict__ sqrt_recipm1,
                                      const float* __restrict__ coef1, const float* __restrict__ coef2,
                                      const float* __restrict__ var, float* __restrict__ x_prev,
                                      float* __restrict__ pred_x0, int n_per_image, size_t total) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int64_t idx = index[i / n_per_image];
  const float xv = x[i];
  // same operation order as the reference (utils/sampler.py:160-164,150-153,200-203)
  const float x0 = sqrt_recip[idx] * xv - sqrt_recipm1[idx] * eps[i];
  const float mean = coef1[idx] * x0 + coef2[idx] * xv;
  const float nz = idx != 0 ? 1.f : 0.f;
  x_prev[i] = mean + nz * sqrtf(var[idx]) * noise[i];
  if (pred_x0 != nullptr) pred_x0[i] = x0;
}

// Partial gaussian blend of latent tiles: out[b, c, y, x] = sum over the given tiles that cover (y, x) of
// weight[y - hi, x - wi] * tiles[t, b, c, y - hi, x - wi]   (numerator of make_tiled_fn, utils/common.py:414-424).
__global__ void tile_blend_kernel(const float* __restrict__ tiles, const int* __restrict__ coords, int ntiles,
                                  const float* __restrict__ weight, float* __restrict__ out, int BC, int H, int W,
                                  int th, int tw, size_t total) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % W);
  const int y = static_cast<int>((i / W) % H);
  const int bc = static_cast<int>(i / (static_cast<size_t>(W) * H));
  float acc = 0.f;
  for (int t = 0; t < ntiles; ++t) {
    const int hi = __ldg(coords + 2 * t), wi = __ldg(coords + 2 * t + 1);
    const int ty = y - hi, tx = x - wi;
    if (ty >= 0 && ty < th && tx >= 0 && tx < tw)
      acc += __ldg(weight + ty * tw + tx) * __ldg(tiles + ((static_cast<size_t>(t) * BC + bc) * th + ty) * tw + tx);
  }
  out[i] = acc;
}


// One level of the wavelet colour fix (utils/common.py:99-147): low = 3x3 binomial blur with dilation `radius` and
// replicate padding (clamped coordinates) of every plane of a [planes, H, W] fp32 tensor.
//   mode 0: out = low                                   (style image, levels 0..3)
//   mode 1: out = low, high (+)= in - low               (content image; `first` level stores instead of adding)
//   mode 2: out = high + low                            (last style level: content_high + style_low)
__global__ void wavelet_level_kernel(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ high,
                                     int H, int W, int radius, int mode, int first, size_t total) {
  PdlScope pdl_scope;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % W);
  const int y = static_cast<int>((i / W) % H);
  const float* plane = in + (i - static_cast<size_t>(y) * W - x);
  const int ym = max(y - radius, 0), yp = min(y + radius, H - 1);
  const int xm = max(x - radius, 0), xp = min(x + radius, W - 1);
  const float* r0 = plane + static_cast<size_t>(ym) * W;
  const float* r1 = plane + static_cast<size_t>(y) * W;
  const float* r2 = plane + static_cast<size_t>(yp) * W;
  // same tap order as a row-major 3x3 correlation
  float low = 0.0625f * __ldg(r0 + xm);
  low = fmaf(0.125f, __ldg(r0 + x), low);
  low = fmaf(0.0625f, __ldg(r0 + xp), low);
  low = fmaf(0.125f, __ldg(r1 + xm), low);
  const float centre = __ldg(r1 + x);
  low = fmaf(0.25f, centre, low);
  low = fmaf(0.125f, __ldg(r1 + xp), low);
  low = fmaf(0.0625f, __ldg(r2 + xm), low);
  low = fmaf(0.125f, __ldg(r2 + x), low);
  low = fmaf(0.0625f, __ldg(r2 + xp), low);
  if (mode == 2) {
    out[i] = high[i] + low;
    return;
  }
  out[i] = low;
  if (mode == 1) high[i] = first ? (centre - low) : (high[i] + (centre - low));
}

static inline unsigned blocks_for(size_t n, int threads) {
  return static_cast<unsigned>((n + threads - 1) / threads);
}

}  // namespace edtr

using namespace edtr;

extern "C" int edtr_nchw_f32_to_nhwc_bf16(const float* X, void* Y, int ldy, int coff, int B, int C, int HW,
                                          void* stream) {
  EDTR_REQUIRE(X && Y && B > 0 && C > 0 && HW > 0 && coff >= 0 && ldy >= coff + C, "bad layout-convert arguments");
  const size_t total = static_cast<size_t>(B) * HW;
  EDTR_LAUNCH(nchw_to_nhwc_kernel, blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      X, reinterpret_cast<__nv_bfloat16*>(Y), ldy, coff, C, HW, total);
  return check_launch("nchw_to_nhwc_kernel");
}

extern "C" int edtr_pointwise_nchw_f32_to_nhwc_bf16(const float* X, const float* Wm, const float* bias, float scale,
                                                    void* Y, int ldy, int coff, int B, int Cin, int Cout, int HW,
                                                    void* stream) {
  EDTR_REQUIRE(X && Wm && Y && B > 0 && HW > 0 && coff >= 0, "bad pointwise-conv arguments");
  EDTR_REQUIRE(Cin > 0 && Cin <= 16 && Cout > 0 && Cout <= 16, "pointwise conv supports 1..16 channels");
  EDTR_REQUIRE(ldy >= coff + Cout, "ldy too small");
  const size_t total = static_cast<size_t>(B) * HW;
  EDTR_LAUNCH(pointwise_nchw_kernel, blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      X, Wm, bias, scale, reinterpret_cast<__nv_bfloat16*>(Y), ldy, coff, Cin, Cout, HW, total);
  return check_launch("pointwise_nchw_kernel");
}

extern "C" int edtr_nhwc_bf16_to_nchw(const void* X, int ldx, void* Y, int B, int C, int HW, int out_f32,
                                      void* stream) {
  EDTR_REQUIRE(X && Y && B > 0 && C > 0 && HW > 0 && ldx >= C, "bad layout-convert arguments");
  EDTR_REQUIRE(B <= 65535 && (C + 31) / 32 <= 65535, "grid too large");
  dim3 grid((HW + 31) / 32, (C + 31) / 32, B), block(32, 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_f32)
    EDTR_LAUNCH((nhwc_to_nchw_kernel<float>), grid, block, 0, st, reinterpret_cast<const __nv_bfloat16*>(X), ldx,
                                                       reinterpret_cast<float*>(Y), C, HW);
  else
    EDTR_LAUNCH((nhwc_to_nchw_kernel<__nv_bfloat16>), grid, block, 0, st, reinterpret_cast<const __nv_bfloat16*>(X), ldx,
                                                               reinterpret_cast<__nv_bfloat16*>(Y), C, HW);
  return check_launch("nhwc_to_nchw_kernel");
}

extern "C" int edtr_cast_f32_to_bf16(const float* X, void* Y, size_t n, void* stream) {
  EDTR_REQUIRE(X && Y, "X/Y is NULL");
  if (n == 0) return EDTR_OK;
  EDTR_LAUNCH(cast_f32_bf16_kernel, blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      X, reinterpret_cast<__nv_bfloat16*>(Y), n);
  return check_launch("cast_f32_bf16_kernel");
}

extern "C" int edtr_upsample2x_bf16(const void* X, int ldx, void* Y, int ldy, int B, int H, int W, int C,
                                    void* stream) {
  EDTR_REQUIRE(X && Y && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad upsample arguments");
  EDTR_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C, "bad strides");
  EDTR_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y)) & 15) == 0, "16-byte alignment");
  const int vpr = C / 8;
  const size_t total = static_cast<size_t>(B) * 4 * H * W * vpr;
  EDTR_LAUNCH(upsample2x_kernel, blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(X), ldx, reinterpret_cast<__nv_bfloat16*>(Y), ldy, H, W, vpr, total);
  return check_launch("upsample2x_kernel");
}

extern "C" int edtr_im2col_bf16(const void* X, int ldx, void* Y, int B, int H, int W, int C, int KH, int KW,
                                int stride, int pad_top, int pad_left, int Ho, int Wo, void* stream) {
  EDTR_REQUIRE(X && Y && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad im2col arguments");
  EDTR_REQUIRE(KH > 0 && KW > 0 && stride > 0 && Ho > 0 && Wo > 0, "bad im2col geometry");
  EDTR_REQUIRE(ldx % 8 == 0 && ldx >= C, "bad stride");
  EDTR_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y)) & 15) == 0, "16-byte alignment");
  const int vpr = C / 8;
  const size_t total = static_cast<size_t>(B) * Ho * Wo * KH * KW * vpr;
  EDTR_LAUNCH(im2col_kernel, blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(X), ldx, reinterpret_cast<__nv_bfloat16*>(Y), H, W, vpr, KH, KW,
      stride, pad_top, pad_left, Ho, Wo, total);
  return check_launch("im2col_kernel");
}

extern "C" int edtr_tile_blend(const float* tiles, const int32_t* coords, int ntiles, const float* weight, float* out,
                               int BC, int H, int W, int th, int tw, void* stream) {
  EDTR_REQUIRE(tiles && coords && weight && out, "NULL argument");
  EDTR_REQUIRE(ntiles > 0 && BC > 0 && H > 0 && W > 0 && th > 0 && tw > 0 && th <= H && tw <= W, "bad tile-blend shape");
  const size_t total = static_cast<size_t>(BC) * H * W;
  EDTR_LAUNCH(tile_blend_kernel, blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), tiles, coords, ntiles,
              weight, out, BC, H, W, th, tw, total);
  return check_launch("tile_blend_kernel");
}

extern "C" int edtr_timestep_embedding(const int64_t* t, void* Y, int B, int dim, float max_period,
                                       void* stream) {
  EDTR_REQUIRE(t && Y && B > 0 && dim >= 2, "bad timestep-embedding arguments");
  EDTR_LAUNCH(timestep_embedding_kernel, B, 128, 0, static_cast<cudaStream_t>(stream), 
      t, reinterpret_cast<__nv_bfloat16*>(Y), dim, logf(max_period));
  return check_launch("timestep_embedding_kernel");
}

extern "C" int edtr_sampler_update(const float* x, const float* eps, const float* noise, const int64_t* index,
                                   const float* sqrt_recip, const float* sqrt_recipm1, const float* coef1,
                                   const float* coef2, const float* var, float* x_prev, float* pred_x0, int B,
                                   int n_per_image, void* stream) {
  EDTR_REQUIRE(x && eps && noise && index && sqrt_recip && sqrt_recipm1 && coef1 && coef2 && var && x_prev,
               "NULL argument");
  EDTR_REQUIRE(B > 0 && n_per_image > 0, "bad sampler-update shape");
  const size_t total = static_cast<size_t>(B) * n_per_image;
  EDTR_LAUNCH(sampler_update_kernel, blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      x, eps, noise, index, sqrt_recip, sqrt_recipm1, coef1, coef2, var, x_prev, pred_x0, n_per_image, total);
  return check_launch("sampler_update_kernel");
}

extern "C" int edtr_wavelet_level(const float* in, float* out, float* high, int planes, int H, int W, int radius,
                                  int mode, int first, void* stream) {
  EDTR_REQUIRE(in && out, "in/out is NULL");
  EDTR_REQUIRE(in != out, "the blur cannot run in place");
  EDTR_REQUIRE(planes > 0 && H > 0 && W > 0 && radius > 0, "bad wavelet geometry");
  EDTR_REQUIRE(mode >= 0 && mode <= 2, "bad mode %d", mode);
  EDTR_REQUIRE(mode == 0 || high != nullptr, "high is NULL");
  const size_t total = static_cast<size_t>(planes) * H * W;
  EDTR_LAUNCH(wavelet_level_kernel, blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), in, out, high,
              H, W, radius, mode, first, total);
  return check_launch("wavelet_level_kernel");
}
