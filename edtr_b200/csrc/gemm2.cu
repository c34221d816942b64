// Persistent CTA-pair (cta_group::2) tcgen05 implicit-GEMM for sm_100a — the main dense kernel.
//
// A cluster of two CTAs (one TPC) owns 256 output rows x `bn` output columns per tile:
//   * each CTA TMA-loads its own 128 rows of A (a 2-D box for GEMMs, a shifted 4-D box per
//     filter tap for 3x3 convolutions; out-of-image pixels are zero-filled = the padding) and
//     its own half (bn/2 rows) of the weight tile; all loads signal the leader CTA's barrier;
//   * the leader issues tcgen05.mma.cta_group::2 (M=256, N=bn, K=16) reading both CTAs'
//     shared memory; accumulators (128 lanes x bn fp32 columns per CTA) are double-buffered
//     in TMEM so the epilogue of tile i overlaps the main loop of tile i+1;
//   * four epilogue warps per CTA drain TMEM in 64-column chunks: optional residual chunk is
//     TMA-loaded into a swizzled staging buffer, bias / time-embedding row vector / residual /
//     SiLU / GEGLU are applied in fp32, the bf16 result is written back to the staging buffer
//     and TMA-stored (fully coalesced, clipped at the matrix edge).
// Tiles are scheduled round-robin over the persistent clusters (grid = 2 x min(tiles, 74)).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace edtr {

constexpr int k2BM = 128;   // rows per CTA (pair = 256)
constexpr int k2BK = 64;
constexpr int k2EpiWarps = 8;                         // two warps per TMEM lane group: even / odd 64-column chunks
constexpr int k2Threads = 64 + 32 * k2EpiWarps;
constexpr int k2MaxStages = 10;
constexpr int k2MaxBN = 256;
constexpr int k2ABytes = k2BM * k2BK * 2;            // 16 KB
constexpr int k2DataBytes = 224 * 1024;              // operand ring + epilogue staging
constexpr int k2WarpStageBytes = 32 * 64 * 2;        // one epilogue warp's staging buffer: 32 rows x 64 bf16 = 4 KB
constexpr int k2MaxRing = 4;                         // staging buffers per epilogue warp (2 or 4)
constexpr int k2SmemBytes = k2DataBytes + 1024 + 1024;

struct Gemm2Params {
  int M, N;                // accumulator matrix: rows, columns (= weight rows)
  int num_kblocks;
  int mode;                // 0 gemm, 1 convolution (taps_x x taps_y filter taps starting at offset (tap_dy0, tap_dx0))
  int H, W, cblocks;
  int taps_x, tap_dy0, tap_dx0;
  int up2x;                // 1: rows are low-resolution pixels of one 2x-upsample phase; D is a 4-D map over the
                           //    phase's output pixels (x and y strides of two pixels)
  int tiles_m, tiles_n;    // 256-row pair tiles, column tiles.  Work items are numbered column-tile major (all row
                           // tiles of column tile 0, then of column tile 1, ...): the only narrower tile is the
                           // last column tile, so the round-robin hands the wide (expensive) tiles out first and
                           // every cluster gets the same mix instead of one parity of clusters taking all wide tiles
  int bn_base;             // width of every column tile but the last (multiple of 64, <= 256)
  int b_box_rows;          // rows of the weight TMA box (= bn_base / 2)
  int splits;              // split-K factor (1 = none); partial tiles go to `ws` as fp32 [splits][M][N]
  int kb_per_split;
  float* ws;
  // LayerNorm folded into the GEMM (see EdtrEpilogue): acc' = rstd * (acc - mean * colsum[n])
  const float* ln_stats;
  int ln_parts;
  float ln_inv_c, ln_eps;
  const float* ln_colsum;
  float* row_stats;        // [M][row_parts][2] sums / sums of squares of the stored values (NULL: not wanted);
  int row_parts;           // row_parts = 2 * tiles_n: one pair per (column tile, epilogue warp group)
  int ring;                // epilogue staging buffers per warp: 4 when the epilogue is the bottleneck (short K), else 2
  int stages;              // pipeline depth: (k2DataBytes - staging) / stage_bytes, <= k2MaxStages
  int stage_bytes;         // 16 KB of A + b_box_rows * 128 B of B
  int geglu;
  int has_residual;
  int act;
  float alpha;
  const float* bias;
  const float* rowvec;
  int rowvec_ld;
  int rows_per_group;
  float* nchw_out;         // != NULL: N <= 8 output channels stored straight to an fp32 [B, N, hw] tensor (the eps /
  int nchw_hw;             //          image / moments outputs of the path); one 64-column tile, bias only
  int halo;                // 1: 3x3 convolution with W % 128 == 0 in "halo" mode — one pipeline stage holds the 130-pixel
                           //    row segment (x0-1 .. x0+128) of ONE input row and channel block plus the three weight
                           //    k-blocks of that row's taps; the three dx taps read it through A descriptors whose start
                           //    address is shifted by one pixel (128 B): a third of the operand traffic of nine boxes
  int a_bytes;             // bytes reserved for A per stage (16 KB; 17 KB in halo mode)
  int b_tile_bytes;        // one weight k-block of this CTA: b_box_rows * 128 B
  int stage_tx;            // bytes the TMA loads of one stage deliver per CTA (expect_tx)
  float* gn_partial;       // != NULL: GroupNorm partial sums of the stored values, [image][gn_slabs][N/4][2] per
  int gn_hw, gn_slabs, gn_slab0;   // (32-row slab, gn_unit-column unit) — see EdtrEpilogue::gn_partial
  int gn_unit;                     // 4 or 2 columns per unit
};

// GroupNorm statistics in the epilogue: v[32] = one row x 32 consecutive stored columns of this lane.  Per 4-column unit
// the lane forms (sum, sum of squares); the 16 values are then summed over the warp's 32 rows with a transposing
// butterfly (at every step a lane keeps one half of its values and hands the other half to its partner), which costs
// 16 shuffles instead of 16 x 5.  Afterwards lane L holds value L >> 1 (unit L >> 2, sum / sum of squares by bit 1) of the
// whole 32-row slab; even lanes store 16 consecutive floats.  Fixed order: deterministic.
__device__ __forceinline__ void gn_partial_store(const float (&v)[32], int lane, float* dst) {
  float a[16];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    a[2 * q] = (v[4 * q] + v[4 * q + 1]) + (v[4 * q + 2] + v[4 * q + 3]);
    a[2 * q + 1] = fmaf(v[4 * q], v[4 * q], fmaf(v[4 * q + 1], v[4 * q + 1], fmaf(v[4 * q + 2], v[4 * q + 2], v[4 * q + 3] * v[4 * q + 3])));
  }
#pragma unroll
  for (int n = 8, bit = 16; n >= 1; n >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? a[i] : a[i + n];
      const float keep = up ? a[i + n] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
  if ((lane & 1) == 0) dst[lane >> 1] = a[0];
}
// Same with 2-column units (group widths that are not multiples of 4: 320 / 32 = 10, 960 / 32 = 30): 32 values, five
// transposing steps, lane L ends up with value L (unit L >> 1, sum / sum of squares by bit 0); 32 consecutive floats.
__device__ __forceinline__ void gn_partial_store2(const float (&v)[32], int lane, float* dst) {
  float a[32];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    a[2 * q] = v[2 * q] + v[2 * q + 1];
    a[2 * q + 1] = fmaf(v[2 * q], v[2 * q], v[2 * q + 1] * v[2 * q + 1]);
  }
#pragma unroll
  for (int n = 16, bit = 16; n >= 1; n >>= 1, bit >>= 1) {
    const bool up = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = up ? a[i] : a[i + n];
      const float keep = up ? a[i + n] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  dst[lane] = a[0];
}

__device__ __forceinline__ float gelu_fast(float x) {
  // x * Phi(x), erf form (model/attention.py:25-27 uses the exact erf GELU).  erf from Abramowitz-Stegun 7.1.27,
  // erf(z) = 1 - (1 + a1 z + a2 z^2 + a3 z^3 + a4 z^4)^-4, |abs err| < 5e-4: one MUFU (rcp) and 11 FMA-pipe
  // instructions — the GEGLU epilogue is instruction-bound (profiles/r02g), and 0.5 |x| 5e-4 is below the bf16
  // resolution of the stored product everywhere.
  const float z = fabsf(x) * 0.70710678118654752f;
  float p = fmaf(0.078108f, z, 0.000972f);
  p = fmaf(p, z, 0.230389f);
  p = fmaf(p, z, 0.278393f);
  p = fmaf(p, z, 1.f);
  p = p * p;
  p = p * p;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));
  const float erfv = copysignf(1.f - r, x);
  const float hx = 0.5f * x;
  return fmaf(hx, erfv, hx);
}

// Packed fp32 pairs (sm_100 FFMA2 / FMUL2): one FMA-pipe instruction per two elements — the GEGLU epilogue is bound by
// instruction issue, so its per-element arithmetic runs on pairs.
__device__ __forceinline__ uint64_t f2pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// (x0 * gelu(g0), x1 * gelu(g1)) with x = accx * s + ax, g = accg * s + ag; gelu as gelu_fast, evaluated on the pair.
__device__ __forceinline__ void geglu_pair(float accx0, float accx1, float accg0, float accg1, uint64_t s2, float ax0,
                                           float ax1, float ag0, float ag1, float& y0, float& y1) {
  const uint64_t x = f2fma(f2pack(accx0, accx1), s2, f2pack(ax0, ax1));
  const uint64_t gt = f2fma(f2pack(accg0, accg1), s2, f2pack(ag0, ag1));
  const uint64_t z = f2mul(gt & 0x7fffffff7fffffffULL, f2pack(0.70710678118654752f, 0.70710678118654752f));   // |g| / sqrt 2
  uint64_t pl = f2fma(f2pack(0.078108f, 0.078108f), z, f2pack(0.000972f, 0.000972f));
  pl = f2fma(pl, z, f2pack(0.230389f, 0.230389f));
  pl = f2fma(pl, z, f2pack(0.278393f, 0.278393f));
  pl = f2fma(pl, z, f2pack(1.f, 1.f));
  pl = f2mul(pl, pl);
  pl = f2mul(pl, pl);
  float p0, p1, r0, r1;
  f2unpack(pl, p0, p1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(p0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(p1));
  // erf(|z|) = 1 - r, sign restored from the gate: xor the gate's sign bits into (1 - r)
  const uint64_t e = f2fma(f2pack(r0, r1), f2pack(-1.f, -1.f), f2pack(1.f, 1.f)) ^ (gt & 0x8000000080000000ULL);
  const uint64_t hg = f2mul(gt, f2pack(0.5f, 0.5f));
  const uint64_t ge = f2fma(hg, e, hg);                       // gelu(g) = 0.5 g (1 + erf)
  f2unpack(f2mul(x, ge), y0, y1);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(k2Threads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD,
             const __grid_constant__ CUtensorMap tmH, const Gemm2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sPipe = smem;                             // stage s: A at s*stage_bytes, B right after it
  uint8_t* sC = sPipe + (k2DataBytes - k2EpiWarps * p.ring * k2WarpStageBytes);   // [warps][ring] x 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + k2DataBytes);
  uint64_t* full_bar = bars;                         // [stages]  (leader's copy is the live one)
  uint64_t* empty_bar = bars + k2MaxStages;          // [stages]
  uint64_t* tfull_bar = bars + 2 * k2MaxStages;      // [2] accumulator ready (multicast to both CTAs)
  uint64_t* tempty_bar = bars + 2 * k2MaxStages + 2; // [2] accumulator drained (leader's copy, 8 arrivals)
  uint64_t* res_bar = bars + 2 * k2MaxStages + 4;    // [4 warps][k2MaxRing] residual chunk landed
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * k2MaxStages + 4 + k2EpiWarps * k2MaxRing);
  const int nstages = p.stages;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int num_work = p.tiles_m * p.tiles_n * p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(p.halo ? &tmH : &tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    if (p.has_residual) tma_prefetch_desc(&tmC);
    for (int s = 0; s < nstages; ++s) {
      mbar_init(&full_bar[s], 2);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 2 * k2EpiWarps);
    }
    for (int b = 0; b < k2EpiWarps * k2MaxRing; ++b) mbar_init(&res_bar[b], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_holder, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();  // PDL: everything above overlapped the previous kernel; global memory is touched only below

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    // The whole warp walks the loop with warp-uniform state (stage counter, tap / channel-block counters: no
    // divisions per k-block) so the addresses live in uniform registers; one elected lane issues.
    const uint32_t pipe_u32 = smem_u32(sPipe);
    const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[0]), 0);
    const uint32_t empty_u32 = smem_u32(&empty_bar[0]);
    uint32_t s = 0, ph = 0;
    for (int w = cluster_id; w < num_work; w += num_clusters) {
      const int tile = w / p.splits, ks = w - tile * p.splits;
      const int kb0 = ks * p.kb_per_split, kb1 = min(p.num_kblocks, kb0 + p.kb_per_split);
      const int tn = tile / p.tiles_m, tm = tile - tn * p.tiles_m;   // column-tile major: see Gemm2Params::tiles_m
      const int m0 = tm * (2 * k2BM) + static_cast<int>(rank) * k2BM;
      const int n_tile0 = tn * p.bn_base;
      const int bn = min(p.bn_base, ((p.N - n_tile0) + 63) & ~63);
      const int nrow0 = n_tile0 + static_cast<int>(rank) * (bn >> 1);
      int cx = 0, cy = 0, cn = 0, cb = 0, tx = 0, ty = 0;
      if (p.mode == 1) {
        cx = (p.W >= k2BM) ? (m0 % p.W) : 0;
        cy = (m0 / p.W) % p.H;
        cn = m0 / (p.W * p.H);
        const int tap0 = kb0 / p.cblocks;
        cb = kb0 - tap0 * p.cblocks;
        ty = tap0 / p.taps_x;
        tx = tap0 - ty * p.taps_x;
        cx += p.tap_dx0;
        cy += p.tap_dy0;
      }
      if (p.halo) {
        // one stage = (input row y + ty - 1, channel block cb): 130 pixels x 64 channels + the weights of taps (ty, 0..2)
        const int hx = m0 % p.W - 1;
        for (int it = 0; it < 3 * p.cblocks; ++it) {
          mbar_wait_u32(empty_u32 + s * 8, ph ^ 1);
          if (elect_one()) {
            const uint32_t fb = full_leader + s * 8;
            const uint32_t sa = pipe_u32 + s * p.stage_bytes;
            mbar_arrive_expect_tx_cluster(fb, p.stage_tx);
            tma_load_4d_pair_u32(sa, &tmH, fb, cb * k2BK, hx, cy + ty, cn);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
              tma_load_2d_pair_u32(sa + p.a_bytes + dx * p.b_tile_bytes, &tmB, fb,
                                   ((ty * 3 + dx) * p.cblocks + cb) * k2BK, nrow0);
          }
          __syncwarp();
          if (++cb == p.cblocks) { cb = 0; ++ty; }
          if (++s == static_cast<uint32_t>(nstages)) { s = 0; ph ^= 1; }
        }
        continue;
      }
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait_u32(empty_u32 + s * 8, ph ^ 1);
        if (elect_one()) {
          const uint32_t fb = full_leader + s * 8;
          const uint32_t sa = pipe_u32 + s * p.stage_bytes;
          mbar_arrive_expect_tx_cluster(fb, p.stage_bytes);
          if (p.mode == 0) {
            tma_load_2d_pair_u32(sa, &tmA, fb, kb * k2BK, m0);
          } else {
            tma_load_4d_pair_u32(sa, &tmA, fb, cb * k2BK, cx + tx, cy + ty, cn);
          }
          tma_load_2d_pair_u32(sa + k2ABytes, &tmB, fb, kb * k2BK, nrow0);
        }
        __syncwarp();
        if (++cb == p.cblocks) {
          cb = 0;
          if (++tx == p.taps_x) { tx = 0; ++ty; }
        }
        if (++s == static_cast<uint32_t>(nstages)) { s = 0; ph ^= 1; }
      }
    }
    // PDL trigger, late on purpose: every operand load of this CTA has been issued, so the next kernel may be
    // scheduled now and run its prologue under this kernel's last MMAs / epilogue.  (A trigger at kernel entry
    // lets waiting dependents grab registers / warp slots that a multi-wave predecessor still needs.)
    pdl_launch_dependents();
  } else if (warp == 1) {
    // ------------------------------------------------------------- UMMA issuer (leader only)
    if (leader) {
      const uint32_t pipe_u32 = smem_u32(sPipe);
      const uint32_t full_u32 = smem_u32(&full_bar[0]);
      const uint32_t empty_u32 = smem_u32(&empty_bar[0]);
      const uint32_t tfull_u32 = smem_u32(&tfull_bar[0]);
      const uint32_t tempty_u32 = smem_u32(&tempty_bar[0]);
      uint32_t s = 0, ph = 0, ti = 0;
      for (int w = cluster_id; w < num_work; w += num_clusters, ++ti) {
        const int tile = w / p.splits, ks = w - tile * p.splits;
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.num_kblocks, kb0 + p.kb_per_split);
        const int tn = tile / p.tiles_m;
        const int n_tile0 = tn * p.bn_base;
        const int bn = min(p.bn_base, ((p.N - n_tile0) + 63) & ~63);
        const uint32_t idesc = umma_idesc_bf16(2 * k2BM, bn);
        const uint32_t a = ti & 1, aph = (ti >> 1) & 1;
        mbar_wait_u32(tempty_u32 + a * 8, aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * k2MaxBN;
        if (p.halo) {
          for (int it = 0; it < 3 * p.cblocks; ++it) {
            mbar_wait_u32(full_u32 + s * 8, ph);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sa = pipe_u32 + s * p.stage_bytes;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                // tap dx reads pixels dx .. dx + 127 of the 130-pixel row segment: start address + dx * 128 B.  The 128 B
                // swizzle of TMA and UMMA is a function of the absolute shared-memory address bits, so the shifted start
                // needs nothing else (the descriptor's matrix-base-offset field stays 0; setting it to dx was measured
                // on B200 and gives wrong results).
                const uint64_t adesc = umma_smem_desc_sw128(sa + dx * 128);
                const uint64_t bdesc = umma_smem_desc_sw128(sa + p.a_bytes + dx * p.b_tile_bytes);
#pragma unroll
                for (int k = 0; k < k2BK / 16; ++k)
                  umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | dx | k) != 0);
              }
              umma_commit_pair_u32(empty_u32 + s * 8, 0x3);
            }
            __syncwarp();
            if (++s == static_cast<uint32_t>(nstages)) { s = 0; ph ^= 1; }
          }
          if (elect_one()) umma_commit_pair_u32(tfull_u32 + a * 8, 0x3);
          __syncwarp();
          continue;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_u32(full_u32 + s * 8, ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = pipe_u32 + s * p.stage_bytes;
            const uint64_t adesc = umma_smem_desc_sw128(sa);
            const uint64_t bdesc = umma_smem_desc_sw128(sa + k2ABytes);
            umma_ss_pair(d_tmem, adesc, bdesc, idesc, kb > kb0);
#pragma unroll
            for (int k = 1; k < k2BK / 16; ++k) umma_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
            umma_commit_pair_u32(empty_u32 + s * 8, 0x3);
          }
          __syncwarp();
          if (++s == static_cast<uint32_t>(nstages)) { s = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit_pair_u32(tfull_u32 + a * 8, 0x3);
        __syncwarp();
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue (both CTAs)
    // Every epilogue warp is independent: it owns TMEM lanes / output rows [lg*32, lg*32+32), a ring
    // of `ring` 4 KB staging buffers and its own TMA traffic (lane 0 issues), so there is no CTA-wide
    // barrier in the epilogue.  Chunks (64 output columns) are numbered globally across tiles (g);
    // the residual chunk g+D is requested while chunk g is processed (D = ring/2 chunks of lead).
    const int lg = warp & 3;
    const int rl = lane;                         // row inside this warp's 32-row slab
    const int ew = warp - 2;
    const int grp = ew >> 2;                     // this warp handles chunks c with (c & 1) == grp
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const int R = p.ring, D = R >> 1, Pend = R - 1 - D;
    uint8_t* wstage = sC + ew * R * k2WarpStageBytes;
    uint64_t* wres = res_bar + ew * k2MaxRing;
    const uint32_t tempty_leader = mapa_u32(smem_u32(&tempty_bar[0]), 0);
    const int n_out = p.geglu ? (p.N >> 1) : p.N;
    const int row_off = static_cast<int>(rank) * k2BM + lg * 32;

    // (work item, chunk) -> coordinates; shared by the processing loop and the residual look-ahead
    auto tile_m0 = [&](int w) { return ((w / p.splits) % p.tiles_m) * (2 * k2BM) + row_off; };
    auto tile_n0 = [&](int w) { return ((w / p.splits) / p.tiles_m) * p.bn_base; };
    auto tile_bn = [&](int w) { return min(p.bn_base, ((p.N - tile_n0(w)) + 63) & ~63); };
    auto tile_chunks = [&](int w) { return (p.geglu ? (tile_bn(w) >> 1) : tile_bn(w)) >> 6; };
    auto tile_outcol0 = [&](int w) { return p.geglu ? (tile_n0(w) >> 1) : tile_n0(w); };

    // look-ahead cursor (lane 0 only): tile coordinates are recomputed once per tile, not per chunk
    int lw = cluster_id, lc = grp, l_m0 = 0, l_col0 = 0, l_chunks = 0;
    uint32_t lgc = 0;              // global index of the chunk the cursor points at
    auto cursor_load_tile = [&]() {   // skips tiles in which this warp owns no chunk
      while (lw < num_work) {
        l_chunks = tile_chunks(lw);
        if (lc < l_chunks) break;
        lw += num_clusters;
      }
      if (lw < num_work) {
        l_m0 = tile_m0(lw);
        l_col0 = tile_outcol0(lw);
      }
    };
    auto issue_residual = [&]() {  // request chunk (lw, lc) into its ring slot and advance the cursor
      if (lw < num_work) {
        const int slot = lgc % R;
        mbar_arrive_expect_tx(&wres[slot], k2WarpStageBytes);
        tma_load_2d(wstage + slot * k2WarpStageBytes, &tmC, &wres[slot], l_col0 + lc * 64, l_m0);
        ++lgc;
        lc += 2;
        if (lc >= l_chunks) {
          lc = grp;
          lw += num_clusters;
          cursor_load_tile();
        }
      }
    };
    if (p.has_residual && lane == 0 && p.splits == 1) {
      cursor_load_tile();
      for (int i = 0; i < D; ++i) issue_residual();
    }

    uint32_t ti = 0, g = 0;
    for (int w = cluster_id; w < num_work; w += num_clusters, ++ti) {
      const int ks = w % p.splits;
      const int m0 = tile_m0(w);
      const int row = m0 + rl;
      const int n_tile0 = tile_n0(w);
      const int bn = tile_bn(w);
      const uint32_t a = ti & 1, aph = (ti >> 1) & 1;
      const uint32_t tacc = trow + a * k2MaxBN;
      // LayerNorm folded into the GEMM: per-row mean / rstd from the producer's partial sums (one lane = one row).
      // The loads are issued before the accumulator wait, four at a time, so their latency hides under the main loop.
      float ln_a = p.alpha, ln_b = 0.f;   // v = acc * ln_a + ln_b * colsum[n] + bias[n]
      if (p.ln_stats != nullptr) {
        // The epilogue is the critical path of these short-K GEMMs, so an L2 round trip per tile would be fully exposed:
        // the statistics of the NEXT tile's row are pulled into L1 now (fire and forget) and are L1 hits one tile later.
        {
          const int wn = w + num_clusters;
          if (wn < num_work) {
            const int rown = tile_m0(wn) + rl;
            if (rown < p.M) {
              const char* pf = reinterpret_cast<const char*>(p.ln_stats) + static_cast<size_t>(rown) * p.ln_parts * 8;
              asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
              asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + p.ln_parts * 8 - 8));
            }
          }
        }
        float s1 = 0.f, s2 = 0.f;
        if (row < p.M) {
          const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + static_cast<size_t>(row) * p.ln_parts;
          int q = 0;
          for (; q + 4 <= p.ln_parts; q += 4) {
            const float2 t0 = __ldg(st + q), t1 = __ldg(st + q + 1), t2 = __ldg(st + q + 2), t3 = __ldg(st + q + 3);
            s1 += (t0.x + t1.x) + (t2.x + t3.x);
            s2 += (t0.y + t1.y) + (t2.y + t3.y);
          }
          for (; q < p.ln_parts; ++q) {
            const float2 t = __ldg(st + q);
            s1 += t.x;
            s2 += t.y;
          }
        }
        const float mu = s1 * p.ln_inv_c;
        const float rstd = rsqrtf(fmaxf(s2 * p.ln_inv_c - mu * mu, 0.f) + p.ln_eps);
        ln_a = p.alpha * rstd;
        ln_b = -ln_a * mu;
      }
      mbar_wait(&tfull_bar[a], aph);
      tc_fence_after();
      if (p.splits > 1) {
        // split-K: raw fp32 partial sums to the workspace; edtr::splitk_reduce_kernel finishes the epilogue
        float* wrow = p.ws + (static_cast<size_t>(ks) * p.M + row) * p.N + n_tile0;
#pragma unroll 1
        for (int c = grp; c < (bn >> 5); c += 2) {
          uint32_t r0[32];
          tmem_ld32(tacc + c * 32, r0);
          tmem_ld_wait();
          if (row < p.M) {
            float4* dst = reinterpret_cast<float4*>(wrow + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst[j] = make_float4(__uint_as_float(r0[4 * j]), __uint_as_float(r0[4 * j + 1]),
                                   __uint_as_float(r0[4 * j + 2]), __uint_as_float(r0[4 * j + 3]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + a * 8);
        continue;
      }
      if (p.nchw_out != nullptr) {
        // few-channel convolution (UNet eps, VAE image / moments): a lane owns one pixel, so for every channel the
        // warp writes 32 consecutive floats of one NCHW plane
        if (grp == 0) {
          uint32_t r0[8];
          tmem_ld8(tacc, r0);
          tmem_ld_wait();
          if (row < p.M) {
            const int img = row / p.nchw_hw;
            float* dst = p.nchw_out + static_cast<size_t>(img) * p.N * p.nchw_hw + (row - img * p.nchw_hw);
#pragma unroll
            for (int c = 0; c < 8; ++c)
              if (c < p.N)
                dst[static_cast<size_t>(c) * p.nchw_hw] =
                    fmaf(__uint_as_float(r0[c]), p.alpha, p.bias != nullptr ? __ldg(p.bias + c) : 0.f);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + a * 8);
        continue;
      }
      const int nchunks = tile_chunks(w);
      if (grp >= nchunks) {   // no chunk of this tile belongs to this warp: release the accumulator at once
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + a * 8);
        if (p.row_stats != nullptr && row < p.M)
          reinterpret_cast<float2*>(p.row_stats)[static_cast<size_t>(row) * p.row_parts + 2 * (n_tile0 / p.bn_base) + grp] =
              make_float2(0.f, 0.f);
        continue;
      }
      const int out_col_tile0 = tile_outcol0(w);
      const int half = bn >> 1;
      const float* rowvec_row = (p.rowvec != nullptr && row < p.M)
                                    ? p.rowvec + static_cast<size_t>(row / p.rows_per_group) * p.rowvec_ld
                                    : nullptr;
      float rs1 = 0.f, rs2 = 0.f;   // producer side of the folded LayerNorm: this row's sums over the warp's chunks
      float* gn_dst = nullptr;      // GroupNorm partial sums of this warp's 32-row slab (warp-uniform)
      if (p.gn_partial != nullptr && m0 < p.M) {
        const int img = m0 / p.gn_hw;
        const int slab = img * p.gn_slabs + p.gn_slab0 + ((m0 - img * p.gn_hw) >> 5);
        gn_dst = p.gn_unit == 2 ? p.gn_partial + static_cast<size_t>(slab) * p.N + n_tile0               // N/2 units x 2
                                : p.gn_partial + (static_cast<size_t>(slab) * (p.N >> 2) + (n_tile0 >> 2)) * 2;
      }
#pragma unroll 1
      for (int c = grp; c < nchunks; c += 2, ++g) {
        const int out_col0 = out_col_tile0 + c * 64;
        const int slot = g % R;
        uint8_t* my_row = wstage + slot * k2WarpStageBytes + rl * 128;
        if (lane == 0) {
          // stores up to chunk g-1-Pend have released their buffers: slot of chunk g+D (and of g) is free
          if (Pend == 0) bulk_wait_group_read<0>(); else bulk_wait_group_read<1>();
          if (p.has_residual) issue_residual();
        }
        __syncwarp();
        if (p.has_residual) mbar_wait(&wres[slot], (g / R) & 1);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {            // 32 columns per TMEM load; rolled to keep the code in the I-cache
          float v[32];
          const int cb = c * 64 + h * 32;        // accumulator / bias column inside the tile
          if (!p.geglu) {
            // The epilogue is the critical path of the short-K GEMMs and only two warps per scheduler hide latency, so
            // everything that does not depend on the accumulator is gathered FIRST, straight into v[] (no extra
            // registers), while the asynchronous TMEM load is in flight: bias, the folded-LayerNorm column term, the
            // time-embedding row vector and the residual chunk.  After the wait one FMA per element remains.
            uint32_t rv[32];
            tmem_ld32(tacc + cb, rv);
            if (p.bias != nullptr) {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + n_tile0 + cb);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(b4 + j);
                v[4 * j] = b.x; v[4 * j + 1] = b.y; v[4 * j + 2] = b.z; v[4 * j + 3] = b.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (p.ln_stats != nullptr) {
              const float4* c4 = reinterpret_cast<const float4*>(p.ln_colsum + n_tile0 + cb);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 t = __ldg(c4 + j);
                v[4 * j] = fmaf(ln_b, t.x, v[4 * j]); v[4 * j + 1] = fmaf(ln_b, t.y, v[4 * j + 1]);
                v[4 * j + 2] = fmaf(ln_b, t.z, v[4 * j + 2]); v[4 * j + 3] = fmaf(ln_b, t.w, v[4 * j + 3]);
              }
            }
            if (rowvec_row != nullptr) {
              const float4* r4 = reinterpret_cast<const float4*>(rowvec_row + out_col0 + h * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(r4 + j);
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
              }
            }
            if (p.has_residual) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 u = *reinterpret_cast<const uint4*>(my_row + (((h * 4 + q) ^ (rl & 7)) << 4));
                const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
                v[8 * q] += f0.x; v[8 * q + 1] += f0.y; v[8 * q + 2] += f1.x; v[8 * q + 3] += f1.y;
                v[8 * q + 4] += f2.x; v[8 * q + 5] += f2.y; v[8 * q + 6] += f3.x; v[8 * q + 7] += f3.y;
              }
            }
            tmem_ld_wait();
            {
              const uint64_t ln_a2 = f2pack(ln_a, ln_a);
#pragma unroll
              for (int j = 0; j < 16; ++j)
                f2unpack(f2fma(f2pack(__uint_as_float(rv[2 * j]), __uint_as_float(rv[2 * j + 1])), ln_a2,
                               f2pack(v[2 * j], v[2 * j + 1])), v[2 * j], v[2 * j + 1]);
            }
          } else {
            uint32_t rv[32], rg[32];
            tmem_ld32(tacc + cb, rv);
            tmem_ld32(tacc + half + cb, rg);   // gate columns
            tmem_ld_wait();
            const float4* bx = reinterpret_cast<const float4*>(p.bias + n_tile0 + cb);
            const float4* bg = reinterpret_cast<const float4*>(p.bias + n_tile0 + half + cb);
            const float4* cx = reinterpret_cast<const float4*>(p.ln_colsum + n_tile0 + cb);
            const float4* cg = reinterpret_cast<const float4*>(p.ln_colsum + n_tile0 + half + cb);
            const uint64_t ln_a2 = f2pack(ln_a, ln_a);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = x4;
              if (p.bias != nullptr) { x4 = __ldg(bx + j); g4 = __ldg(bg + j); }
              if (p.ln_stats != nullptr) {
                const float4 sx = __ldg(cx + j), sg = __ldg(cg + j);
                x4.x = fmaf(ln_b, sx.x, x4.x); x4.y = fmaf(ln_b, sx.y, x4.y);
                x4.z = fmaf(ln_b, sx.z, x4.z); x4.w = fmaf(ln_b, sx.w, x4.w);
                g4.x = fmaf(ln_b, sg.x, g4.x); g4.y = fmaf(ln_b, sg.y, g4.y);
                g4.z = fmaf(ln_b, sg.z, g4.z); g4.w = fmaf(ln_b, sg.w, g4.w);
              }
              geglu_pair(__uint_as_float(rv[4 * j]), __uint_as_float(rv[4 * j + 1]), __uint_as_float(rg[4 * j]),
                         __uint_as_float(rg[4 * j + 1]), ln_a2, x4.x, x4.y, g4.x, g4.y, v[4 * j], v[4 * j + 1]);
              geglu_pair(__uint_as_float(rv[4 * j + 2]), __uint_as_float(rv[4 * j + 3]), __uint_as_float(rg[4 * j + 2]),
                         __uint_as_float(rg[4 * j + 3]), ln_a2, x4.z, x4.w, g4.z, g4.w, v[4 * j + 2], v[4 * j + 3]);
            }
            if (rowvec_row != nullptr) {
              const float4* r4 = reinterpret_cast<const float4*>(rowvec_row + out_col0 + h * 32);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 b = __ldg(r4 + j);
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
              }
            }
            if (p.has_residual) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 u = *reinterpret_cast<const uint4*>(my_row + (((h * 4 + q) ^ (rl & 7)) << 4));
                const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
                v[8 * q] += f0.x; v[8 * q + 1] += f0.y; v[8 * q + 2] += f1.x; v[8 * q + 3] += f1.y;
                v[8 * q + 4] += f2.x; v[8 * q + 5] += f2.y; v[8 * q + 6] += f3.x; v[8 * q + 7] += f3.y;
              }
            }
          }
          if (p.act == EDTR_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
          } else if (p.act >= EDTR_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = act_extra_f(v[j], p.act);
          }
          if (p.row_stats != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { rs1 += v[j]; rs2 = fmaf(v[j], v[j], rs2); }
          }
          if (gn_dst != nullptr) {
            if (p.gn_unit == 2) gn_partial_store2(v, lane, gn_dst + cb);
            else gn_partial_store(v, lane, gn_dst + (cb >> 1));
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16(v[8 * q], v[8 * q + 1]); o.y = pack_bf16(v[8 * q + 2], v[8 * q + 3]);
            o.z = pack_bf16(v[8 * q + 4], v[8 * q + 5]); o.w = pack_bf16(v[8 * q + 6], v[8 * q + 7]);
            *reinterpret_cast<uint4*>(my_row + (((h * 4 + q) ^ (rl & 7)) << 4)) = o;
          }
        }
        if (c + 2 >= nchunks) {
          // last TMEM read of this tile by this warp: hand the accumulator stage back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader + a * 8);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (!p.up2x) {
            tma_store_2d(&tmD, wstage + slot * k2WarpStageBytes, out_col0, m0);
          } else {  // 32 consecutive low-resolution pixels -> (x, y, image) of the phase's output grid
            const int ox = (p.W >= 32) ? (m0 % p.W) : 0;
            tma_store_4d(&tmD, wstage + slot * k2WarpStageBytes, out_col0, ox, (m0 / p.W) % p.H, m0 / (p.W * p.H));
          }
          bulk_commit_group();
        }
      }
      if (p.row_stats != nullptr && row < p.M)   // one pair per (row, column tile, warp group): 2 * tiles_n per row
        reinterpret_cast<float2*>(p.row_stats)[static_cast<size_t>(row) * p.row_parts + 2 * (n_tile0 / p.bn_base) + grp] =
            make_float2(rs1, rs2);
    }
    if (lane == 0) bulk_wait_group<0>();
    (void)n_out;
  }
  __syncwarp();  // role branches diverge inside a warp; the cluster barrier below is .aligned
  pdl_launch_dependents();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// Sums the split-K partial tiles in a fixed order and applies the fused epilogue (bias / row vector /
// residual / SiLU), 8 columns per thread.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, int M, int N, float alpha,
                     const float* __restrict__ bias, const float* __restrict__ rowvec, int rowvec_ld,
                     int rows_per_group, const __nv_bfloat16* residual, int ldr, __nv_bfloat16* out, int ldc,
                     int act) {
  pdl_wait();  // PDL: wait for the previous kernel's results (the trigger is at the end of the body)
  const int vpr = N >> 3;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<size_t>(M) * vpr) return;
  const int row = static_cast<int>(idx / vpr), col = static_cast<int>(idx - static_cast<size_t>(row) * vpr) << 3;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  const size_t plane = static_cast<size_t>(M) * N;
  const float* src = ws + static_cast<size_t>(row) * N + col;
  for (int s = 0; s < splits; ++s) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + s * plane));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + s * plane + 4));
    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] *= alpha;
  if (bias != nullptr) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(bias + col));
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
  }
  if (rowvec != nullptr) {
    const float* rv = rowvec + static_cast<size_t>(row / rows_per_group) * rowvec_ld + col;
    const float4 a = __ldg(reinterpret_cast<const float4*>(rv));
    const float4 b = __ldg(reinterpret_cast<const float4*>(rv + 4));
    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
  }
  if (residual != nullptr) {
    const uint4 u = *reinterpret_cast<const uint4*>(residual + static_cast<size_t>(row) * ldr + col);
    const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y), f2 = unpack_bf16(u.z), f3 = unpack_bf16(u.w);
    v[0] += f0.x; v[1] += f0.y; v[2] += f1.x; v[3] += f1.y; v[4] += f2.x; v[5] += f2.y; v[6] += f3.x; v[7] += f3.y;
  }
  if (act == EDTR_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = silu_f(v[i]);
  } else if (act >= EDTR_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = act_extra_f(v[i], act);
  }
  uint4 o;
  o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * ldc + col) = o;
  pdl_launch_dependents();
}

int prime_gemm2_attributes() {
  cudaError_t e = cudaFuncSetAttribute(gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemBytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(gemm2): %s", cudaGetErrorString(e));
    return EDTR_ERR_CUDA;
  }
  return EDTR_OK;
}

// Column tiling and split-K.  Every column tile but the last is `bn_base` wide (a multiple of 64,
// <= 256).  Width and split factor are chosen with a small cost model: waves over the 74 CTA pairs x
// (k-blocks per work item x cycles per k-block + a fixed fill/drain cost), where a k-block of a bn-wide
// pair tile costs max(2*bn, 300) cycles (narrow tiles are bound by operand fetch, not by the tensor
// pipe) and splitting pays for the fp32 round trip + reduce launch.  GEGLU is pinned to 256 columns,
// unsplit, because its weight interleave is done at pack time.
static void plan_tiles(int M, int N, int nkb, int geglu, size_t ws_bytes, int max_clusters, int* tiles_n, int* bn_base,
                       int* splits) {
  *splits = 1;
  if (geglu) {
    *tiles_n = N / k2MaxBN;
    *bn_base = k2MaxBN;
    return;
  }
  const int tiles_m = (M + 2 * k2BM - 1) / (2 * k2BM);
  long best_cost = -1;
  int best_t = 1, best_b = 64, best_s = 1;
  const size_t split_max = ws_bytes / (static_cast<size_t>(M) * N * sizeof(float));
  const int max_split_ws = split_max > 16 ? 16 : static_cast<int>(split_max);
  for (int cap = k2MaxBN; cap >= 64; cap -= 64) {
    const int t = (N + cap - 1) / cap;
    int base = ((N + t - 1) / t + 63) & ~63;
    if (base > cap) base = cap;
    const int tn = (N + base - 1) / base;
    const long per_kb = 2 * base > 300 ? 2 * base : 300;
    for (int sp = 1; sp <= 16; ++sp) {
      if (sp > 1 && (sp > max_split_ws || nkb / sp < 8)) break;
      const int kbs = (nkb + sp - 1) / sp;
      if (sp > 1 && kbs * (sp - 1) >= nkb) continue;  // an empty split
      const long waves = (static_cast<long>(tiles_m) * tn * sp + max_clusters - 1) / max_clusters;
      const long cost = waves * (kbs * per_kb + 3000) + (sp > 1 ? 10000 : 0);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_t = tn;
        best_b = base;
        best_s = sp;
      }
    }
  }
  *tiles_n = best_t;
  *bn_base = best_b;
  *splits = best_s;
}

bool gemm2_disabled();

bool gemm2_eligible(int M, int N, const EdtrEpilogue* ep) {
  if (gemm2_disabled()) return false;
  static const bool nchw_ok = [] {
    const char* e = getenv("EDTR_GEMM2_NCHW");   // debugging switch: 0 sends the few-channel outputs to the single-CTA kernel
    return e == nullptr || e[0] != '0';
  }();
  if (ep->out_mode == EDTR_OUT_NCHW_F32)
    return nchw_ok && M >= 256 && N <= 8 && ep->hw > 0 && M % ep->hw == 0 && ep->residual == nullptr && ep->rowvec == nullptr &&
           ep->act == EDTR_ACT_NONE;
  if (ep->out_mode != EDTR_OUT_BF16) return false;
  const int n_out = ep->act == EDTR_ACT_GEGLU ? N / 2 : N;
  if (ep->act == EDTR_ACT_GEGLU) return N % k2MaxBN == 0;  // any M: the weight interleave is tied to this kernel
  if (M < 256) return false;
  if (n_out % 8 != 0 || N % 64 != 0) return false;
  if (ep->rowvec != nullptr && ((ep->rowvec_ld & 3) != 0 || (reinterpret_cast<uintptr_t>(ep->rowvec) & 15) != 0))
    return false;
  return true;
}

int gemm2_tile_n(int N, int geglu) {
  int t, b, sp;
  plan_tiles(1 << 20, N, 64, geglu, 0, 74, &t, &b, &sp);
  return b;
}

int gemm2_row_stats_parts(int M, int N, int K, const EdtrEpilogue* ep) {
  int max_clusters = ep->max_clusters > 0 ? ep->max_clusters : 74;
  if (max_clusters > 74) max_clusters = 74;
  int t, b, sp;
  plan_tiles(M, N, K / k2BK, 0, 0, max_clusters, &t, &b, &sp);   // launches that write row statistics never split
  return 2 * t;
}

size_t gemm2_workspace_size(int M, int N, int K) {
  if (M < 256 || N % 64 != 0 || K % k2BK != 0) return 0;
  int t, b, sp;
  plan_tiles(M, N, K / k2BK, 0, static_cast<size_t>(-1), 74, &t, &b, &sp);
  if (sp <= 1) return 0;
  return static_cast<size_t>(sp) * M * N * sizeof(float);
}

bool gemm2_disabled() {
  const char* e = getenv("EDTR_GEMM_V1");  // debugging switch: force the single-CTA kernel
  return e != nullptr && e[0] == '1';
}

// A/B tensor maps are built by the caller (gemm or conv geometry); C/D maps are built here.
int launch_gemm2(const CUtensorMap& tmA, const void* Wt, int ldw, int K, int M, int N, int mode, int H, int W,
                 int cblocks, const EdtrEpilogue* ep, cudaStream_t stream, int taps_x, int tap_dy0, int tap_dx0,
                 const CUtensorMap* tmD_up, const CUtensorMap* tmA_halo) {
  Gemm2Params p{};
  p.M = M; p.N = N; p.num_kblocks = K / k2BK; p.mode = mode; p.H = H; p.W = W; p.cblocks = cblocks;
  p.taps_x = taps_x; p.tap_dy0 = tap_dy0; p.tap_dx0 = tap_dx0; p.up2x = tmD_up != nullptr;
  p.geglu = ep->act == EDTR_ACT_GEGLU;
  const bool nchw = ep->out_mode == EDTR_OUT_NCHW_F32;
  int max_clusters = ep->max_clusters > 0 ? ep->max_clusters : 74;
  if (max_clusters > 74) max_clusters = 74;
  // split-K scratch (fp32 partial tiles).  The folded LayerNorm and the row statistics live in the tile epilogue only,
  // so launches that use them do not split (they are N = C GEMMs with K <= 4C: the planner would not split them anyway)
  size_t ws_partial_bytes = 0;
  if (ep->workspace != nullptr && !p.up2x && !nchw && ep->ln_stats == nullptr && ep->row_stats == nullptr &&
      ep->gn_partial == nullptr)
    ws_partial_bytes = static_cast<size_t>(ep->workspace_bytes);
  if (ep->gn_partial != nullptr) {
    if (nchw || p.geglu || ep->gn_hw <= 0 || ep->gn_hw % 32 != 0 || M % ep->gn_hw != 0 || ep->gn_slab0 < 0 ||
        ep->gn_slabs < ep->gn_slab0 + ep->gn_hw / 32 || (reinterpret_cast<uintptr_t>(ep->gn_partial) & 7) != 0) {
      set_error("gn_partial needs a bf16 output, act != GEGLU, gn_hw %% 32 == 0, gn_hw | M, gn_slabs >= gn_slab0 + gn_hw/32 "
                "(gn_hw %d, M %d, gn_slabs %d, gn_slab0 %d)", ep->gn_hw, M, ep->gn_slabs, ep->gn_slab0);
      return EDTR_ERR_INVALID;
    }
    p.gn_partial = ep->gn_partial;
    p.gn_hw = ep->gn_hw;
    p.gn_slabs = ep->gn_slabs;
    p.gn_slab0 = ep->gn_slab0;
    p.gn_unit = ep->gn_unit == 2 ? 2 : 4;
    if (ep->gn_unit != 0 && ep->gn_unit != 2 && ep->gn_unit != 4) {
      set_error("gn_unit must be 0 (= 4), 2 or 4 (got %d)", ep->gn_unit);
      return EDTR_ERR_INVALID;
    }
  }
  plan_tiles(M, N, p.num_kblocks, p.geglu, ws_partial_bytes, max_clusters, &p.tiles_n, &p.bn_base, &p.splits);
  if (nchw) {
    p.nchw_out = static_cast<float*>(ep->out);
    p.nchw_hw = ep->hw;
  }
  p.kb_per_split = (p.num_kblocks + p.splits - 1) / p.splits;
  p.ws = static_cast<float*>(ep->workspace);
  p.ln_stats = ep->ln_stats;
  p.ln_parts = ep->ln_parts;
  p.ln_inv_c = ep->ln_c > 0 ? 1.f / static_cast<float>(ep->ln_c) : 0.f;
  p.ln_eps = ep->ln_eps;
  p.ln_colsum = ep->ln_colsum;
  p.row_stats = ep->row_stats;
  p.row_parts = 2 * p.tiles_n;
  if (ep->row_stats != nullptr && ep->row_stats_cap < p.row_parts) {
    set_error("row_stats_cap %d < %d pairs per row this launch writes (edtr_gemm_row_stats_parts)", ep->row_stats_cap,
              p.row_parts);
    return EDTR_ERR_INVALID;
  }
  p.b_box_rows = p.bn_base / 2;
  p.a_bytes = k2ABytes;
  p.b_tile_bytes = p.b_box_rows * k2BK * 2;
  p.stage_bytes = k2ABytes + p.b_tile_bytes;
  p.stage_tx = p.stage_bytes;
  // Halo mode (see Gemm2Params::halo): narrow convolutions at W % 128 == 0 are bound by the operand traffic of the nine
  // tap boxes through L2 (measured 8 TB/s at 512 x 512 x 128 channels), not by the tensor pipe.  EDTR_CONV_HALO=0
  // switches it off (A/B).
  static const int halo_mode = [] {
    const char* e = getenv("EDTR_CONV_HALO");
    return e == nullptr ? 1 : atoi(e);
  }();
  if (halo_mode > 0 && tmA_halo != nullptr && mode == 1 && taps_x == 3 && !p.up2x && W % (k2BM) == 0 && p.splits == 1 &&
      p.bn_base <= 128 && p.num_kblocks == 9 * cblocks) {
    p.halo = 1;
    p.a_bytes = 17 * 1024;                       // 130 pixels x 128 B, rounded up to the 1024 B swizzle atom
    p.stage_bytes = p.a_bytes + 3 * p.b_tile_bytes;
    p.stage_tx = 130 * k2BK * 2 + 3 * p.b_tile_bytes;
  }
  // two staging buffers per epilogue warp; a ring of four (residual requested two chunks ahead) was measured on
  // B200 and changes nothing: the short-K GEMMs are bound by launch / fill / drain latency, not by the residual
  p.ring = 2;
  p.stages = (k2DataBytes - k2EpiWarps * p.ring * k2WarpStageBytes) / p.stage_bytes;
  if (p.stages > k2MaxStages) p.stages = k2MaxStages;
  p.tiles_m = (M + 2 * k2BM - 1) / (2 * k2BM);
  p.has_residual = ep->residual != nullptr;
  p.act = ep->act; p.alpha = ep->alpha; p.bias = ep->bias; p.rowvec = ep->rowvec; p.rowvec_ld = ep->rowvec_ld;
  p.rows_per_group = ep->rows_per_group > 0 ? ep->rows_per_group : 1;
  const int n_out = p.geglu ? N / 2 : N;
  CUtensorMap tmB, tmC, tmD;
  int rc;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(N)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {k2BK, static_cast<uint32_t>(p.b_box_rows)};
    rc = make_tmap_bf16(&tmB, Wt, 2, dims, strides, box);
    if (rc) return rc;
  }
  if (tmD_up != nullptr) {
    tmD = *tmD_up;
  } else if (nchw) {
    tmD = tmA;   // unused (plain stores); any valid descriptor will do for the prefetch
  } else {
    uint64_t dims[2] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(ep->ldc) * 2};
    uint32_t box[2] = {64, 32};
    rc = make_tmap_bf16(&tmD, ep->out, 2, dims, strides, box);
    if (rc) return rc;
  }
  if (p.has_residual) {
    uint64_t dims[2] = {static_cast<uint64_t>(n_out), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(ep->ldr) * 2};
    uint32_t box[2] = {64, 32};
    rc = make_tmap_bf16(&tmC, ep->residual, 2, dims, strides, box);
    if (rc) return rc;
  } else {
    tmC = tmD;
  }
  const int work = p.tiles_m * p.tiles_n * p.splits;
  const int clusters = work < max_clusters ? work : max_clusters;
  if (p.splits > 1) p.has_residual = 0;  // the reduce kernel adds it
  if (p.up2x && p.splits > 1) {
    set_error("internal: split-K is not available for the up-sampling convolution");
    return EDTR_ERR_INVALID;
  }
  EDTR_LAUNCH(gemm2_kernel, 2 * clusters, k2Threads, k2SmemBytes, stream, tmA, tmB, tmC, tmD,
              p.halo ? *tmA_halo : tmA, p);
  rc = check_launch("gemm2_kernel");
  if (rc || p.splits == 1) return rc;
  const size_t nvec = static_cast<size_t>(M) * (N / 8);
  EDTR_LAUNCH(splitk_reduce_kernel, static_cast<unsigned>((nvec + 255) / 256), 256, 0, stream,
      p.ws, p.splits, M, N, ep->alpha, ep->bias, ep->rowvec, ep->rowvec_ld, p.rows_per_group,
      reinterpret_cast<const __nv_bfloat16*>(ep->residual), ep->ldr, reinterpret_cast<__nv_bfloat16*>(ep->out),
      ep->ldc, ep->act);
  return check_launch("splitk_reduce_kernel");
}

}  // namespace edtr
