// SwinIR pre-restoration network (model/swinir.py): the ops that are not plain GEMM / conv / norm launches.
//
//  * pixel_unshuffle_kernel: fp32 NCHW image -> bf16 channels-last [B, H/r, W/r, C r^2] with the RGB mean removed
//    (model/swinir.py:859-860, PixelUnshuffle in conv_first :700-704); 8 source floats -> one 16-byte store.
//  * window_attention_kernel: W-MSA / SW-MSA of one SwinTransformerBlock (model/swinir.py:120-151, :245-281).  One CTA
//    (4 warps) per (window, head): the cyclic shift, the window partition and their inverses are index arithmetic on
//    the token grid (no roll / permute passes), Q/K/V of the window's 64 tokens are staged in shared memory, each warp
//    owns 16 query rows and evaluates S = Q K^T (mma.sync m16n8k16 bf16, 64 x 64 x 32 per head: far below the size
//    where a tcgen05 / TMEM pipeline pays), adds the relative-position bias and the shift mask, soft-maxes the row in
//    registers (quad shuffles) and multiplies by V.  Heads are 30 wide in the reference; they are stored 32 wide
//    (zero weights in the two pad columns), so a head is one 64-byte run of a token's q / k / v row.
//
// Checked on B200 against a fp32 PyTorch evaluation (tests/test_kernels_gpu.py::test_window_attention) and inside the
// SwinIR engine against the live-reference fixture (tests/test_engine_gpu.py); 36 us per block at 8 x 64 x 64 tokens.
#include "common.cuh"
#include "ptx.cuh"

namespace edtr {

__global__ void __launch_bounds__(256)
pixel_unshuffle_kernel(const float* __restrict__ X, __nv_bfloat16* __restrict__ Y, int ldy, int C, int H, int W, int r,
                       float m0, float m1, float m2, float scale, size_t total) {
  PdlScope pdl_scope;
  // one thread per (b, y, x, c, dy): r == 8 consecutive source pixels of one row -> 8 consecutive output channels
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int Ho = H / r, Wo = W / r;
  size_t t = idx;
  const int dy = static_cast<int>(t % r); t /= r;
  const int c = static_cast<int>(t % C); t /= C;
  const int x = static_cast<int>(t % Wo); t /= Wo;
  const int y = static_cast<int>(t % Ho); t /= Ho;
  const int b = static_cast<int>(t);
  const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
  const float* src = X + ((static_cast<size_t>(b) * C + c) * H + (y * r + dy)) * W + x * r;
  const float4 a = __ldg(reinterpret_cast<const float4*>(src));
  const float4 d = __ldg(reinterpret_cast<const float4*>(src + 4));
  uint4 o;
  o.x = pack_bf16((a.x - mean) * scale, (a.y - mean) * scale);
  o.y = pack_bf16((a.z - mean) * scale, (a.w - mean) * scale);
  o.z = pack_bf16((d.x - mean) * scale, (d.y - mean) * scale);
  o.w = pack_bf16((d.z - mean) * scale, (d.w - mean) * scale);
  __nv_bfloat16* dst = Y + ((static_cast<size_t>(b) * Ho + y) * Wo + x) * ldy + (c * r + dy) * r;
  *reinterpret_cast<uint4*>(dst) = o;
}

// D (16x8, fp32) += A (16x16, bf16, row) * B (16x8, bf16, col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kWin = 8;                 // window side (tokens)
constexpr int kWinTok = kWin * kWin;    // 64 tokens per window
constexpr int kHd = 32;                 // stored head width (30 real + 2 zero pads)
constexpr int kRowPad = 40;             // shared-memory row pitch in bf16 (80 B: conflict-free fragment loads)

__global__ void __launch_bounds__(128)
window_attention_kernel(const __nv_bfloat16* __restrict__ QKV, int ld, __nv_bfloat16* __restrict__ O, int ldo,
                        int H, int W, int heads, int shift, float scale, const float* __restrict__ bias,
                        const float* __restrict__ mask) {
  PdlScope pdl_scope;
  __shared__ __align__(16) __nv_bfloat16 sQ[kWinTok][kRowPad];
  __shared__ __align__(16) __nv_bfloat16 sK[kWinTok][kRowPad];
  __shared__ __align__(16) __nv_bfloat16 sV[kWinTok][kRowPad];
  __shared__ int tok[kWinTok];   // global token (row of QKV / O) of every window position

  const int head = blockIdx.y;
  const int wpi = (H / kWin) * (W / kWin);         // windows per image
  const int img = blockIdx.x / wpi;
  const int win = blockIdx.x - img * wpi;          // window index inside the (shifted) image = mask index
  const int wy = win / (W / kWin), wx = win - wy * (W / kWin);
  const int tid = threadIdx.x;
  if (tid < kWinTok) {
    // position (iy, ix) of the window in the rolled image is source pixel ((y + shift) mod H, (x + shift) mod W):
    // torch.roll(x, (-shift, -shift)) (model/swinir.py:255-257); the inverse roll writes back to the same pixel
    const int iy = tid >> 3, ix = tid & 7;
    int y = wy * kWin + iy + shift, x = wx * kWin + ix + shift;
    if (y >= H) y -= H;
    if (x >= W) x -= W;
    tok[tid] = (img * H + y) * W + x;
  }
  __syncthreads();
  const int C3 = heads * kHd;                      // q | k | v blocks of the fused projection are C3 columns apart
  for (int i = tid; i < kWinTok * 4 * 3; i += blockDim.x) {   // 64 tokens x 4 vectors x {q, k, v}
    const int m = i / (kWinTok * 4);
    const int r = (i >> 2) & (kWinTok - 1), v = i & 3;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(QKV + static_cast<size_t>(tok[r]) * ld + m * C3 + head * kHd + v * 8));
    __nv_bfloat16* dst = m == 0 ? &sQ[r][v * 8] : (m == 1 ? &sK[r][v * 8] : &sV[r][v * 8]);
    *reinterpret_cast<uint4*>(dst) = u;
  }
  __syncthreads();

  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = warp * 16;                        // this warp's 16 query rows
  // S[16 x 64] = Q[16 x 32] K^T: 8 key tiles of 8, 2 k-steps of 16
  float s[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    uint32_t a[4];
    a[0] = *reinterpret_cast<const uint32_t*>(&sQ[q0 + g][kk * 16 + 2 * t]);
    a[1] = *reinterpret_cast<const uint32_t*>(&sQ[q0 + g + 8][kk * 16 + 2 * t]);
    a[2] = *reinterpret_cast<const uint32_t*>(&sQ[q0 + g][kk * 16 + 8 + 2 * t]);
    a[3] = *reinterpret_cast<const uint32_t*>(&sQ[q0 + g + 8][kk * 16 + 8 + 2 * t]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&sK[j * 8 + g][kk * 16 + 2 * t]);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&sK[j * 8 + g][kk * 16 + 8 + 2 * t]);
      mma_bf16_16816(s[j], a, b0, b1);
    }
  }
  // logits = scale * S + bias[head] (+ mask[window]); rows g and g + 8 of the warp's slab, columns 8 j + 2 t + {0, 1}
  const float* bh = bias + (static_cast<size_t>(head) * kWinTok + q0) * kWinTok;
  const float* mw = mask != nullptr ? mask + (static_cast<size_t>(win) * kWinTok + q0) * kWinTok : nullptr;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int col = j * 8 + 2 * t;
    float2 b0 = __ldg(reinterpret_cast<const float2*>(bh + g * kWinTok + col));
    float2 b1 = __ldg(reinterpret_cast<const float2*>(bh + (g + 8) * kWinTok + col));
    if (mw != nullptr) {
      const float2 m0 = __ldg(reinterpret_cast<const float2*>(mw + g * kWinTok + col));
      const float2 m1 = __ldg(reinterpret_cast<const float2*>(mw + (g + 8) * kWinTok + col));
      b0.x += m0.x; b0.y += m0.y; b1.x += m1.x; b1.y += m1.y;
    }
    s[j][0] = fmaf(s[j][0], scale, b0.x);
    s[j][1] = fmaf(s[j][1], scale, b0.y);
    s[j][2] = fmaf(s[j][2], scale, b1.x);
    s[j][3] = fmaf(s[j][3], scale, b1.y);
    mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
    mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
  }
  // a row lives in the four lanes of a quad
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s[j][0] = __expf(s[j][0] - mx0);
    s[j][1] = __expf(s[j][1] - mx0);
    s[j][2] = __expf(s[j][2] - mx1);
    s[j][3] = __expf(s[j][3] - mx1);
    l0 += s[j][0] + s[j][1];
    l1 += s[j][2] + s[j][3];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // O[16 x 32] = P[16 x 64] V[64 x 32]: 4 d tiles of 8, 4 k-steps of 16 keys; the S accumulator tiles 2 kk and 2 kk + 1
  // are exactly the A fragment of k-step kk
  float o[4][4];
#pragma unroll
  for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      // B[k = key][n = d]: keys 16 kk + 2 t + {0, 1} (+ 8), column 8 n + g
      const int k0 = kk * 16 + 2 * t, d = n * 8 + g;
      const uint32_t b0 = static_cast<uint32_t>(__bfloat16_as_ushort(sV[k0][d])) |
                          (static_cast<uint32_t>(__bfloat16_as_ushort(sV[k0 + 1][d])) << 16);
      const uint32_t b1 = static_cast<uint32_t>(__bfloat16_as_ushort(sV[k0 + 8][d])) |
                          (static_cast<uint32_t>(__bfloat16_as_ushort(sV[k0 + 9][d])) << 16);
      mma_bf16_16816(o[n], a, b0, b1);
    }
  }
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  __nv_bfloat16* r0 = O + static_cast<size_t>(tok[q0 + g]) * ldo + head * kHd;
  __nv_bfloat16* r1 = O + static_cast<size_t>(tok[q0 + g + 8]) * ldo + head * kHd;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    *reinterpret_cast<uint32_t*>(r0 + n * 8 + 2 * t) = pack_bf16(o[n][0] * i0, o[n][1] * i0);
    *reinterpret_cast<uint32_t*>(r1 + n * 8 + 2 * t) = pack_bf16(o[n][2] * i1, o[n][3] * i1);
  }
}

}  // namespace edtr

using namespace edtr;

extern "C" int edtr_pixel_unshuffle_f32_to_nhwc_bf16(const float* X, void* Y, int ldy, int B, int C, int H, int W, int r,
                                                     const float* mean3, float scale, void* stream) {
  EDTR_REQUIRE(X && Y, "X/Y is NULL");
  EDTR_REQUIRE(B > 0 && C > 0 && C <= 3 && H > 0 && W > 0 && r == 8, "pixel unshuffle supports r == 8 and <= 3 channels");
  EDTR_REQUIRE(H % r == 0 && W % r == 0, "H and W must be multiples of %d", r);
  EDTR_REQUIRE(ldy % 8 == 0 && ldy >= C * r * r && (reinterpret_cast<uintptr_t>(Y) & 15) == 0, "bad Y stride / alignment");
  EDTR_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0, "X must be 16-byte aligned");
  const size_t total = static_cast<size_t>(B) * (H / r) * (W / r) * C * r;
  const float m0 = mean3 ? mean3[0] : 0.f, m1 = mean3 ? mean3[1] : 0.f, m2 = mean3 ? mean3[2] : 0.f;   // host pointer
  EDTR_LAUNCH(pixel_unshuffle_kernel, static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream),
              X, reinterpret_cast<__nv_bfloat16*>(Y), ldy, C, H, W, r, m0, m1, m2, scale, total);
  return check_launch("pixel_unshuffle_kernel");
}

extern "C" int edtr_window_attention_bf16(const void* QKV, int ld, void* O, int ldo, int B, int H, int W, int heads,
                                          int shift, float scale, const float* bias, const float* mask, void* stream) {
  EDTR_REQUIRE(QKV && O && bias, "QKV/O/bias is NULL");
  EDTR_REQUIRE(B > 0 && heads > 0 && H % kWin == 0 && W % kWin == 0 && H > 0 && W > 0, "token grid must be a multiple of 8x8");
  EDTR_REQUIRE(shift >= 0 && shift < kWin && (shift == 0 || mask != nullptr || (H == kWin && W == kWin)),
               "a shifted block needs its attention mask");
  EDTR_REQUIRE(ld % 8 == 0 && ld >= 3 * heads * kHd && ldo % 8 == 0 && ldo >= heads * kHd, "row strides too small");
  EDTR_REQUIRE(((reinterpret_cast<uintptr_t>(QKV) | reinterpret_cast<uintptr_t>(O)) & 15) == 0, "QKV/O must be 16-byte aligned");
  EDTR_REQUIRE(((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(mask)) & 7) == 0, "bias/mask must be 8-byte aligned");
  const long long windows = static_cast<long long>(B) * (H / kWin) * (W / kWin);
  EDTR_REQUIRE(windows < (1ll << 31) && heads <= 65535, "grid too large");
  dim3 grid(static_cast<unsigned>(windows), heads, 1);
  EDTR_LAUNCH(window_attention_kernel, grid, 128, 0, static_cast<cudaStream_t>(stream),
              reinterpret_cast<const __nv_bfloat16*>(QKV), ld, reinterpret_cast<__nv_bfloat16*>(O), ldo, H, W, heads, shift,
              scale, bias, mask);
  return check_launch("window_attention_kernel");
}
