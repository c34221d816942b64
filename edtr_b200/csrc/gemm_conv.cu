// tcgen05 / TMEM / TMA implicit-GEMM kernel for sm_100a.
//
// One kernel template serves every dense contraction on the ControlLDM path:
//   mode 0: out = A[M,K] * Wt[N,K]^T                    (nn.Linear, 1x1 conv)
//   mode 1: 3x3 / stride-1 / pad-1 convolution over a channels-last activation
//           [B,H,W,Cin]; the A tile of every filter tap is a shifted 4-D TMA box,
//           out-of-bounds pixels are zero-filled by the TMA unit (= the padding).
// Tile: 128 output rows x BN output columns, K in blocks of 64 bf16 (one 128-byte
// swizzle row).  Warp roles: warp 0 = TMA producer, warp 1 = UMMA issuer (+TMEM
// allocation), warps 2..5 = epilogue (TMEM -> registers -> fused bias / time-emb
// / residual / SiLU / GEGLU -> global).
#include <stdarg.h>

#include "common.cuh"
#include "ptx.cuh"

namespace edtr {

constexpr int kBM = 128;          // rows per tile (UMMA M)
constexpr int kBK = 64;           // K elements per stage (128 B of bf16)
constexpr int kGemmThreads = 192; // 6 warps

struct GemmKernelParams {
  int M, N;
  int num_kblocks;
  int mode;             // 0 gemm, 1 conv3x3
  int H, W, cblocks;    // conv geometry (cblocks = Cin / 64)
  EdtrEpilogue ep;
};

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN <= 160) ? 3 : 4;
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kTmemCols = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int kSmemBytes = kStages * (kABytes + kBBytes) + 1024 /*align*/ + 256 /*bars*/;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  return act == EDTR_ACT_SILU ? silu_f(v) : act_extra_f(v, act);
}

// Adds bias / rowvec / residual to 32 accumulator columns of one row.
__device__ __forceinline__ void epilogue_addends(float (&v)[32], const EdtrEpilogue& ep, int row,
                                                 int col0, int ncols_valid) {
  if (ep.bias != nullptr) {
    if (ncols_valid == 32) {
      const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 b = __ldg(b4 + j);
        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    } else {
      for (int j = 0; j < ncols_valid; ++j) v[j] += __ldg(ep.bias + col0 + j);
    }
  }
  if (ep.rowvec != nullptr) {
    const float* rv = ep.rowvec + static_cast<size_t>(row / ep.rows_per_group) * ep.rowvec_ld + col0;
    if (ncols_valid == 32 && (ep.rowvec_ld & 3) == 0) {
      const float4* r4 = reinterpret_cast<const float4*>(rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 b = __ldg(r4 + j);
        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    } else {
      for (int j = 0; j < ncols_valid; ++j) v[j] += __ldg(rv + j);
    }
  }
  if (ep.residual != nullptr) {
    const __nv_bfloat16* res =
        reinterpret_cast<const __nv_bfloat16*>(ep.residual) + static_cast<size_t>(row) * ep.ldr + col0;
    if (ncols_valid == 32) {
      const uint4* r4 = reinterpret_cast<const uint4*>(res);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u = r4[j];
        float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
        v[8 * j + 0] += a.x; v[8 * j + 1] += a.y; v[8 * j + 2] += b.x; v[8 * j + 3] += b.y;
        v[8 * j + 4] += c.x; v[8 * j + 5] += c.y; v[8 * j + 6] += d.x; v[8 * j + 7] += d.y;
      }
    } else {
      for (int j = 0; j < ncols_valid; ++j) v[j] += __bfloat162float(res[j]);
    }
  }
}

// Stores 32 finished columns of one row. n_out = number of columns of the stored matrix.
__device__ __forceinline__ void epilogue_store(const float (&v)[32], const EdtrEpilogue& ep, int row,
                                               int col0, int ncols_valid, int n_out) {
  if (ep.out_mode == EDTR_OUT_BF16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + static_cast<size_t>(row) * ep.ldc + col0;
    if (ncols_valid == 32) {
      uint4* o4 = reinterpret_cast<uint4*>(o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
        u.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
        u.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
        o4[j] = u;
      }
    } else {
      for (int j = 0; j < ncols_valid; ++j) o[j] = __float2bfloat16(v[j]);
    }
  } else if (ep.out_mode == EDTR_OUT_F32) {
    float* o = reinterpret_cast<float*>(ep.out) + static_cast<size_t>(row) * ep.ldc + col0;
    if (ncols_valid == 32 && (ep.ldc & 3) == 0) {
      float4* o4 = reinterpret_cast<float4*>(o);
#pragma unroll
      for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      for (int j = 0; j < ncols_valid; ++j) o[j] = v[j];
    }
  } else {
    const int img = row / ep.hw, pix = row - img * ep.hw;
    const size_t base = (static_cast<size_t>(img) * n_out + col0) * ep.hw + pix;
    if (ep.out_mode == EDTR_OUT_NCHW_F32) {
      float* o = reinterpret_cast<float*>(ep.out) + base;
#pragma unroll 8
      for (int j = 0; j < 32; ++j)
        if (j < ncols_valid) o[static_cast<size_t>(j) * ep.hw] = v[j];
    } else {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(ep.out) + base;
#pragma unroll 8
      for (int j = 0; j < 32; ++j)
        if (j < ncols_valid) o[static_cast<size_t>(j) * ep.hw] = __float2bfloat16(v[j]);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(kGemmThreads, (BN <= 160) ? 2 : 1)
gemm_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmKernelParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + Cfg::kStages * Cfg::kBBytes);
  uint64_t* full_bar = bars;                     // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;     // [kStages]
  uint64_t* accum_bar = bars + 2 * Cfg::kStages; // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kBM;
  const int n0 = blockIdx.x * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();  // PDL: everything above overlapped the previous kernel; global memory is touched only below

  if (warp == 0) {
    // ------------------------------------------------------ TMA producer
    // warp-uniform loop state (stage / tap / channel-block counters, no divisions), one elected lane issues
    const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
    const uint32_t full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);
    int cx = 0, cy = 0, cn = 0;
    if (p.mode == 1) {
      cx = ((p.W >= kBM) ? (m0 % p.W) : 0) - 1;
      cy = (m0 / p.W) % p.H - 1;
      cn = m0 / (p.W * p.H);
    }
    uint32_t s = 0, ph = 0;
    int cb = 0, tx = 0, ty = 0;
    for (int kb = 0; kb < p.num_kblocks; ++kb) {
      mbar_wait_u32(empty_u32 + s * 8, ph ^ 1);
      if (elect_one()) {
        const uint32_t fb = full_u32 + s * 8;
        mbar_arrive_expect_tx_u32(fb, Cfg::kABytes + Cfg::kBBytes);
        if (p.mode == 0) {
          tma_load_2d_u32(sA_u32 + s * Cfg::kABytes, &tmA, fb, kb * kBK, m0);
        } else {
          tma_load_4d_u32(sA_u32 + s * Cfg::kABytes, &tmA, fb, cb * kBK, cx + tx, cy + ty, cn);
        }
        tma_load_2d_u32(sB_u32 + s * Cfg::kBBytes, &tmB, fb, kb * kBK, n0);
      }
      __syncwarp();
      if (++cb == p.cblocks) {
        cb = 0;
        if (++tx == 3) { tx = 0; ++ty; }
      }
      if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------- UMMA issuer
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
    const uint32_t sA_u32 = smem_u32(sA), sB_u32 = smem_u32(sB);
    const uint32_t full_u32 = smem_u32(full_bar), empty_u32 = smem_u32(empty_bar);
    uint32_t s = 0, ph = 0;
    for (int kb = 0; kb < p.num_kblocks; ++kb) {
      mbar_wait_u32(full_u32 + s * 8, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = umma_smem_desc_sw128(sA_u32 + s * Cfg::kABytes);
        const uint64_t bdesc = umma_smem_desc_sw128(sB_u32 + s * Cfg::kBBytes);
        // +32 B per 16-element K step inside the 128 B swizzle row (encoded >>4)
        umma_ss(tmem_base, adesc, bdesc, idesc, kb != 0);
#pragma unroll
        for (int k = 1; k < kBK / 16; ++k) umma_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
        umma_commit_u32(empty_u32 + s * 8);  // frees the smem stage when these MMAs retire
      }
      __syncwarp();
      if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
    }
    if (elect_one()) umma_commit(accum_bar);  // accumulator complete
    __syncwarp();
  } else {
    // ---------------------------------------------------------- epilogue
    const int lg = warp & 3;  // TMEM lane group this warp may access
    const int row = m0 + lg * 32 + lane;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const EdtrEpilogue& ep = p.ep;
    float ln_a = ep.alpha, ln_b = 0.f;   // folded LayerNorm: v = acc * ln_a + ln_b * colsum[n] (+ bias ...)
    if (ep.ln_stats != nullptr && row < p.M) {
      const float2* st = reinterpret_cast<const float2*>(ep.ln_stats) + static_cast<size_t>(row) * ep.ln_parts;
      float s1 = 0.f, s2 = 0.f;
      for (int q = 0; q < ep.ln_parts; ++q) {
        const float2 t = __ldg(st + q);
        s1 += t.x;
        s2 += t.y;
      }
      const float inv_c = 1.f / static_cast<float>(ep.ln_c);
      const float mu = s1 * inv_c;
      const float rstd = rsqrtf(fmaxf(s2 * inv_c - mu * mu, 0.f) + ep.ln_eps);
      ln_a = ep.alpha * rstd;
      ln_b = -ln_a * mu;
    }
    if (ep.act != EDTR_ACT_GEGLU) {
      float rs1 = 0.f, rs2 = 0.f;   // row statistics of the stored values: one pair per (row, column tile)
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(trow + c * 32, r);
        tmem_ld_wait();
        if (row < p.M) {
          const int nv = min(32, p.N - col0);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * ln_a;
          if (ep.ln_stats != nullptr) {   // folded LayerNorm (see EdtrEpilogue): N % 32 == 0 is checked on the host
#pragma unroll 8
            for (int j = 0; j < 32; ++j) v[j] = fmaf(ln_b, __ldg(ep.ln_colsum + col0 + j), v[j]);
          }
          epilogue_addends(v, ep, row, col0, nv);
          if (ep.act == EDTR_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
          } else if (ep.act >= EDTR_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = act_extra_f(v[j], ep.act);
          }
          if (ep.row_stats != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nv) { rs1 += v[j]; rs2 = fmaf(v[j], v[j], rs2); }
          }
          epilogue_store(v, ep, row, col0, nv, p.N);
        }
      }
      if (ep.row_stats != nullptr && row < p.M)
        reinterpret_cast<float2*>(ep.row_stats)[static_cast<size_t>(row) * gridDim.x + blockIdx.x] = make_float2(rs1, rs2);
    } else {
      // value columns [0, BN/2) and gate columns [BN/2, BN) of this tile
      constexpr int HALF = BN / 2;
      const int n_out = p.N / 2;
#pragma unroll 1
      for (int c = 0; c < HALF / 32; ++c) {
        uint32_t rx[32], rg[32];
        tmem_ld32(trow + c * 32, rx);
        tmem_ld32(trow + HALF + c * 32, rg);
        tmem_ld_wait();
        const int colx = n0 + c * 32;            // column in the interleaved weight order
        const int col_out = blockIdx.x * HALF + c * 32;
        if (row < p.M && col_out < n_out) {
          const int nv = min(32, n_out - col_out);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(rx[j]) * ep.alpha;
            float g = __uint_as_float(rg[j]) * ep.alpha;
            if (ep.bias != nullptr && j < nv) {
              x += __ldg(ep.bias + colx + j);
              g += __ldg(ep.bias + colx + HALF + j);
            }
            v[j] = x * gelu_erf_f(g);
          }
          EdtrEpilogue ep2 = ep;
          ep2.bias = nullptr;
          epilogue_addends(v, ep2, row, col_out, nv);
          epilogue_store(v, ep, row, col_out, nv, n_out);
        }
      }
    }
    tc_fence_before();
  }
  pdl_launch_dependents();  // late trigger (see gemm2.cu)
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------ host side
static int pick_bn(int M, int N, int act) {
  (void)M;
  if (act == EDTR_ACT_GEGLU) return 128;
  if (N <= 64) return 64;
  if (N % 128 != 0 && N % 160 == 0) return 160;
  return 128;
}

template <int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmKernelParams& p,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_conv_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm<%d>): %s", BN, cudaGetErrorString(e));
      return EDTR_ERR_CUDA;
    }
    attr_set = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + kBM - 1) / kBM, 1);
  EDTR_LAUNCH((gemm_conv_kernel<BN>), grid, kGemmThreads, Cfg::kSmemBytes, stream, tmA, tmB, p);
  return check_launch("gemm_conv_kernel");
}

static int dispatch_gemm(int bn, const CUtensorMap& tmA, const CUtensorMap& tmB,
                         const GemmKernelParams& p, cudaStream_t stream) {
  switch (bn) {
    case 64: return launch_gemm<64>(tmA, tmB, p, stream);
    case 128: return launch_gemm<128>(tmA, tmB, p, stream);
    case 160: return launch_gemm<160>(tmA, tmB, p, stream);
    default: set_error("unsupported tile N %d", bn); return EDTR_ERR_INVALID;
  }
}

int prime_gemm_attributes() {
  cudaError_t e;
  e = cudaFuncSetAttribute(gemm_conv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<64>::kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_conv_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128>::kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_conv_kernel<160>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<160>::kSmemBytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    return EDTR_ERR_CUDA;
  }
  return EDTR_OK;
}

bool gemm2_eligible(int M, int N, const EdtrEpilogue* ep);
int gemm2_tile_n(int N, int geglu);
int gemm2_row_stats_parts(int M, int N, int K, const EdtrEpilogue* ep);
bool gemm2_disabled();
int launch_gemm2(const CUtensorMap& tmA, const void* Wt, int ldw, int K, int M, int N, int mode, int H, int W,
                 int cblocks, const EdtrEpilogue* ep, cudaStream_t stream, int taps_x = 3, int tap_dy0 = -1,
                 int tap_dx0 = -1, const CUtensorMap* tmD_up = nullptr, const CUtensorMap* tmA_halo = nullptr);

static int check_epilogue(const EdtrEpilogue* ep, int M, int N) {
  EDTR_REQUIRE(ep != nullptr && ep->out != nullptr, "epilogue/out is NULL");
  EDTR_REQUIRE(ep->act >= 0 && ep->act <= 5, "bad act %d", ep->act);
  EDTR_REQUIRE(ep->out_mode >= 0 && ep->out_mode <= 3, "bad out_mode %d", ep->out_mode);
  const int n_out = ep->act == EDTR_ACT_GEGLU ? N / 2 : N;
  if (ep->out_mode == EDTR_OUT_BF16) {
    EDTR_REQUIRE(ep->ldc >= n_out, "ldc %d < N %d", ep->ldc, n_out);
    EDTR_REQUIRE(ep->ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(ep->out) & 15) == 0,
                 "bf16 output must be 16-byte aligned (ldc %% 8 == 0)");
  } else if (ep->out_mode == EDTR_OUT_F32) {
    EDTR_REQUIRE(ep->ldc >= n_out, "ldc %d < N %d", ep->ldc, n_out);
  } else {
    EDTR_REQUIRE(ep->hw > 0 && M % ep->hw == 0, "NCHW output needs hw | M (hw %d, M %d)", ep->hw, M);
  }
  if (ep->residual != nullptr)
    EDTR_REQUIRE(ep->ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(ep->residual) & 15) == 0,
                 "residual must be 16-byte aligned (ldr %% 8 == 0)");
  if (ep->workspace != nullptr)
    EDTR_REQUIRE((reinterpret_cast<uintptr_t>(ep->workspace) & 255) == 0, "workspace must be 256-byte aligned");
  EDTR_REQUIRE(ep->max_clusters >= 0, "max_clusters must be >= 0");
  if (ep->ln_stats != nullptr) {
    EDTR_REQUIRE(ep->ln_colsum != nullptr && ep->ln_parts > 0 && ep->ln_c > 0, "folded LayerNorm needs ln_colsum / ln_parts / ln_c");
    EDTR_REQUIRE(N % 32 == 0 && ((reinterpret_cast<uintptr_t>(ep->ln_stats) & 7) == 0) &&
                     ((reinterpret_cast<uintptr_t>(ep->ln_colsum) & 15) == 0),
                 "folded LayerNorm needs N %% 32 == 0, 8-byte aligned ln_stats and 16-byte aligned ln_colsum");
    EDTR_REQUIRE(ep->act != EDTR_ACT_GEGLU || (N % 256 == 0 && !gemm2_disabled()),
                 "folded LayerNorm + GEGLU runs on the CTA-pair kernel only (N %% 256 == 0)");
  }
  if (ep->row_stats != nullptr)
    EDTR_REQUIRE(ep->out_mode == EDTR_OUT_BF16 && ep->act != EDTR_ACT_GEGLU &&
                     (reinterpret_cast<uintptr_t>(ep->row_stats) & 7) == 0,
                 "row_stats needs a bf16 output, act != GEGLU and an 8-byte aligned buffer");
  if (ep->gn_partial != nullptr)
    EDTR_REQUIRE(gemm2_eligible(M, N, ep) && ep->out_mode == EDTR_OUT_BF16 && ep->act != EDTR_ACT_GEGLU,
                 "gn_partial is written by the CTA-pair kernel only (M >= 256, N %% 64 == 0, bf16 output, act != GEGLU)");
  if (ep->rowvec != nullptr)
    EDTR_REQUIRE(ep->rows_per_group > 0 && ep->rowvec_ld >= n_out, "bad rowvec geometry");
  if (ep->bias != nullptr)
    EDTR_REQUIRE((reinterpret_cast<uintptr_t>(ep->bias) & 15) == 0, "bias must be 16-byte aligned");
  if (ep->act == EDTR_ACT_GEGLU) {
    EDTR_REQUIRE(N % 128 == 0, "GEGLU needs N %% 128 == 0 (N %d)", N);
    EDTR_REQUIRE(ep->out_mode == EDTR_OUT_BF16 || N % 256 != 0 || gemm2_disabled(),
                 "GEGLU with N %% 256 == 0 runs on the CTA-pair kernel, which stores bf16 only");
  }
  return EDTR_OK;
}

}  // namespace edtr

using namespace edtr;

extern "C" int edtr_gemm_tile_n(int M, int N, int K, int act) {
  (void)K;
  if (act == EDTR_ACT_GEGLU && N % 256 == 0 && !gemm2_disabled()) return gemm2_tile_n(N, 1);
  return pick_bn(M, N, act);
}

extern "C" int edtr_gemm_row_stats_parts(int M, int N, int K, const EdtrEpilogue* ep) {
  if (ep == nullptr || M <= 0 || N <= 0 || K <= 0 || K % kBK != 0) return 0;
  if (gemm2_eligible(M, N, ep)) return gemm2_row_stats_parts(M, N, K, ep);
  const int bn = pick_bn(M, N, ep->act);
  return (N + bn - 1) / bn;
}

extern "C" int edtr_gemm_bf16(const void* A, int lda, const void* Wt, int ldw, int M, int N, int K,
                              const EdtrEpilogue* ep, void* stream) {
  EDTR_REQUIRE(A && Wt, "A/Wt is NULL");
  EDTR_REQUIRE(M > 0 && N > 0 && K > 0, "bad GEMM shape %dx%dx%d", M, N, K);
  EDTR_REQUIRE(K % kBK == 0, "K (%d) must be a multiple of 64", K);
  EDTR_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, "lda/ldw must be >= K and multiples of 8");
  EDTR_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Wt)) & 15) == 0, "A/Wt must be 16-byte aligned");
  int rc = check_epilogue(ep, M, N);
  if (rc) return rc;
  const int bn = pick_bn(M, N, ep->act);
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(lda) * 2};
    uint32_t box[2] = {kBK, kBM};
    rc = make_tmap_bf16(&tmA, A, 2, dims, strides, box);
    if (rc) return rc;
  }
  if (gemm2_eligible(M, N, ep))
    return launch_gemm2(tmA, Wt, ldw, K, M, N, 0, 0, 0, 0, ep, static_cast<cudaStream_t>(stream));
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {kBK, static_cast<uint32_t>(bn)};
    rc = make_tmap_bf16(&tmB, Wt, 2, dims, strides, box);
    if (rc) return rc;
  }
  GemmKernelParams p{};
  p.M = M; p.N = N; p.num_kblocks = K / kBK; p.mode = 0;
  p.ep = *ep;
  EDTR_REQUIRE(ep->row_stats == nullptr || ep->row_stats_cap >= (N + bn - 1) / bn,
               "row_stats_cap %d < %d pairs per row this launch writes", ep->row_stats_cap, (N + bn - 1) / bn);
  return dispatch_gemm(bn, tmA, tmB, p, static_cast<cudaStream_t>(stream));
}

// 2x nearest up-sampling followed by a 3x3 / pad-1 convolution, evaluated as four 2x2-tap convolutions on the
// low-resolution input (one per output phase (py, px)): out[b, 2y+py, 2x+px] = sum over the 2x2 taps of
// Wp[py][px] * in[b, y+dy, x+dx].  2.25x fewer MACs and operand bytes than convolving the up-sampled tensor.
extern "C" int edtr_conv3x3_up2x_bf16(const void* X, int ldx, int B, int H, int W, int Cin, const void* Wt4,
                                      int Cout, const EdtrEpilogue* ep, void* stream) {
  EDTR_REQUIRE(X && Wt4 && ep && ep->out, "X/Wt4/out is NULL");
  EDTR_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "bad conv shape");
  EDTR_REQUIRE(Cin % kBK == 0 && Cout % 64 == 0, "Cin and Cout must be multiples of 64");
  EDTR_REQUIRE(ldx % 8 == 0 && ldx >= Cin, "ldx must be >= Cin and a multiple of 8");
  EDTR_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Wt4) | reinterpret_cast<uintptr_t>(ep->out)) & 15) == 0,
               "X/Wt4/out must be 16-byte aligned");
  EDTR_REQUIRE(ep->out_mode == EDTR_OUT_BF16 && ep->act != EDTR_ACT_GEGLU && ep->residual == nullptr && ep->rowvec == nullptr,
               "the up-sampling convolution stores bf16 and supports bias + a pointwise activation only");
  EDTR_REQUIRE(ep->ldc % 8 == 0 && ep->ldc >= Cout, "bad ldc");
  EDTR_REQUIRE(W >= 8 && (W & (W - 1)) == 0 && (W >= 128 || 128 % W == 0), "W (%d) must be a power of two >= 8", W);
  int bw, bh, bn_img;
  if (W >= kBM) { bw = kBM; bh = 1; bn_img = 1; }
  else {
    bw = W;
    const int rows = kBM / W;
    if (H >= rows) { EDTR_REQUIRE(H % rows == 0, "H (%d) must be a multiple of %d", H, rows); bh = rows; bn_img = 1; }
    else { EDTR_REQUIRE(rows % H == 0, "H (%d) must divide %d", H, rows); bh = H; bn_img = rows / H; }
  }
  const int M = B * H * W;
  EDTR_REQUIRE(M >= 256 && !gemm2_disabled(), "the up-sampling convolution needs the CTA-pair kernel (M >= 256)");
  CUtensorMap tmA;
  int rc;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(B)};
    uint64_t strides[3] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(ldx) * 2 * W,
                           static_cast<uint64_t>(ldx) * 2 * W * H};
    uint32_t box[4] = {kBK, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bn_img)};
    rc = make_tmap_bf16(&tmA, X, 4, dims, strides, box);
    if (rc) return rc;
  }
  // 32-pixel store slabs of the epilogue warps inside the low-resolution grid
  const int sw = W >= 32 ? 32 : W;
  int sh = 32 / sw, sn = 1;
  if (sh > H) { sn = sh / H; sh = H; }
  const int K = 4 * Cin;
  const size_t ldc = static_cast<size_t>(ep->ldc);
  EdtrEpilogue ep_phase = *ep;   // GroupNorm partial sums: the four phases fill disjoint slab ranges of every image
  if (ep->gn_partial != nullptr) {
    EDTR_REQUIRE((H * W) % 32 == 0 && ep->gn_slabs >= 4 * (H * W / 32), "gn_partial of the up-sampling convolution needs "
                 "H*W %% 32 == 0 and gn_slabs >= 4*H*W/32 (got %d)", ep->gn_slabs);
    ep_phase.gn_hw = H * W;
  }
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      ep_phase.gn_slab0 = ep->gn_partial != nullptr ? (py * 2 + px) * (H * W / 32) : 0;
      CUtensorMap tmD;
      __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(ep->out) + (static_cast<size_t>(py) * 2 * W + px) * ldc;
      uint64_t dims[4] = {static_cast<uint64_t>(Cout), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                          static_cast<uint64_t>(B)};
      uint64_t strides[3] = {2 * ldc * 2, 2 * (2 * static_cast<uint64_t>(W)) * ldc * 2,
                             (2 * static_cast<uint64_t>(H)) * (2 * static_cast<uint64_t>(W)) * ldc * 2};
      uint32_t box[4] = {64, static_cast<uint32_t>(sw), static_cast<uint32_t>(sh), static_cast<uint32_t>(sn)};
      rc = make_tmap_bf16(&tmD, base, 4, dims, strides, box);
      if (rc) return rc;
      const __nv_bfloat16* wp = reinterpret_cast<const __nv_bfloat16*>(Wt4) + static_cast<size_t>(py * 2 + px) * Cout * K;
      rc = launch_gemm2(tmA, wp, K, K, M, Cout, 1, H, W, Cin / kBK, &ep_phase, static_cast<cudaStream_t>(stream), 2,
                        py - 1, px - 1, &tmD);
      if (rc) return rc;
    }
  return EDTR_OK;
}

extern "C" int edtr_conv3x3_bf16(const void* X, int ldx, int B, int H, int W, int Cin, const void* Wt,
                                 int Cout, const EdtrEpilogue* ep, void* stream) {
  EDTR_REQUIRE(X && Wt, "X/Wt is NULL");
  EDTR_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "bad conv shape");
  EDTR_REQUIRE(Cin % kBK == 0, "Cin (%d) must be a multiple of 64", Cin);
  EDTR_REQUIRE(ldx % 8 == 0 && ldx >= Cin, "ldx must be >= Cin and a multiple of 8");
  EDTR_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Wt)) & 15) == 0, "X/Wt must be 16-byte aligned");
  // tile geometry: 128 consecutive output pixels form a (bn x bh x bw) box
  int bw, bh, bn_img;
  if (W >= kBM) {
    EDTR_REQUIRE(W % kBM == 0, "W (%d) must be a multiple of 128 when >= 128", W);
    bw = kBM; bh = 1; bn_img = 1;
  } else {
    EDTR_REQUIRE(kBM % W == 0 && W >= 8, "W (%d) must be a power of two in [8,64]", W);
    bw = W;
    const int rows = kBM / W;
    if (H >= rows) {
      EDTR_REQUIRE(H % rows == 0, "H (%d) must be a multiple of %d", H, rows);
      bh = rows; bn_img = 1;
    } else {
      EDTR_REQUIRE(rows % H == 0, "H (%d) must divide %d", H, rows);
      bh = H; bn_img = rows / H;
    }
  }
  const int M = B * H * W;
  int rc = check_epilogue(ep, M, Cout);
  if (rc) return rc;
  EDTR_REQUIRE(ep->act != EDTR_ACT_GEGLU, "GEGLU is not defined for convolutions");
  const int bn = pick_bn(M, Cout, ep->act);
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                        static_cast<uint64_t>(B)};
    uint64_t strides[3] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(ldx) * 2 * W,
                           static_cast<uint64_t>(ldx) * 2 * W * H};
    uint32_t box[4] = {kBK, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bn_img)};
    rc = make_tmap_bf16(&tmA, X, 4, dims, strides, box);
    if (rc) return rc;
  }
  const int K = 9 * Cin;
  if (gemm2_eligible(M, Cout, ep)) {
    CUtensorMap tmH;
    const bool halo = W % kBM == 0;   // 130-pixel row segments (one halo pixel on either side) for the kernel's halo mode
    if (halo) {
      uint64_t dims[4] = {static_cast<uint64_t>(Cin), static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                          static_cast<uint64_t>(B)};
      uint64_t strides[3] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(ldx) * 2 * W,
                             static_cast<uint64_t>(ldx) * 2 * W * H};
      uint32_t box[4] = {kBK, static_cast<uint32_t>(kBM + 2), 1, 1};
      rc = make_tmap_bf16(&tmH, X, 4, dims, strides, box);
      if (rc) return rc;
    }
    return launch_gemm2(tmA, Wt, K, K, M, Cout, 1, H, W, Cin / kBK, ep, static_cast<cudaStream_t>(stream), 3, -1, -1,
                        nullptr, halo ? &tmH : nullptr);
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(Cout)};
    uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {kBK, static_cast<uint32_t>(bn)};
    rc = make_tmap_bf16(&tmB, Wt, 2, dims, strides, box);
    if (rc) return rc;
  }
  GemmKernelParams p{};
  p.M = M; p.N = Cout; p.num_kblocks = K / kBK; p.mode = 1;
  p.H = H; p.W = W; p.cblocks = Cin / kBK;
  p.ep = *ep;
  return dispatch_gemm(bn, tmA, tmB, p, static_cast<cudaStream_t>(stream));
}
