// HBM-bound normalisation kernels: GroupNorm statistics, GroupNorm apply (+SiLU),
// LayerNorm, row softmax.  All activations are bf16 channels-last; every global
// access is a 16-byte vector of 8 channels, adjacent threads touch adjacent
// vectors (coalesced), reductions use warp shuffles + shared memory.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace edtr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
// SiLU for the normalisation kernels: y * sigmoid(y) with sigmoid(y) = 0.5 + 0.5 tanh(y / 2) — ONE MUFU op (tanh.approx,
// relative error 2^-11, far below the bf16 resolution of the stored result) instead of the two (ex2 + rcp) of the exact
// form.  The GroupNorm + SiLU launches are bound by instruction issue / MUFU, not by memory (profiles/r02q).
__device__ __forceinline__ float silu_tanh(float y) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * y));
  return y * fmaf(0.5f, t, 0.5f);
}
// y = x * sc + sf (+ SiLU) on one 16-byte vector of 8 channels.
__device__ __forceinline__ uint4 norm8(const uint4& u, const float (&sc)[8], const float (&sf)[8], bool silu) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  float f[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], sc[i], sf[i]);
  if (silu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = silu_tanh(f[i]);
  }
  uint4 o;
  o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
  return o;
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16(f[0], f[1]); u.y = pack_bf16(f[2], f[3]);
  u.z = pack_bf16(f[4], f[5]); u.w = pack_bf16(f[6], f[7]);
  return u;
}

// ------------------------------------------------------------ GroupNorm stats
// Deterministic two-level reduction, no atomics.  grid (pixel chunks, B, channel slabs);
// block = vslab * ppar threads (vslab = 16-byte vectors per slab).  Thread (v, q) owns vector v
// and walks pixels q, q+ppar, ... U at a time (U independent 16-byte loads in flight: 8 for the L2-resident tensors
// of the UNet, whose launches are latency-bound, 4 for the HBM-streaming VAE tensors that run 8 CTAs per SM),
// keeping its 8 per-channel sums in registers.  The CTA folds the ppar partial rows through
// shared memory in a fixed order and writes per-(chunk, channel-group) partial sums to
// partial[b][chunk][group][2]; slabs that split a group write their share into separate
// slab slots: partial[b][chunk][slab][group][2].
template <int U>
__global__ void __launch_bounds__(256)
groupnorm_stats_kernel(const __nv_bfloat16* __restrict__ X, int ldx, int HW, int C, int groups,
                       int rows_per_cta, int vslab, float* __restrict__ partial) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  extern __shared__ float sh[];  // [ppar][2][cslab]
  const int cslab = vslab * 8;
  const int ppar = blockDim.x / vslab;
  const int v = threadIdx.x % vslab;
  const int q = threadIdx.x / vslab;
  const int b = blockIdx.y;
  const int c0 = blockIdx.z * cslab;
  const int p0 = blockIdx.x * rows_per_cta;
  const int p1 = min(HW, p0 + rows_per_cta);
  float s[8], ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = ss[i] = 0.f;
  if (q < ppar) {
    const size_t xstep = static_cast<size_t>(ppar) * ldx;
    const __nv_bfloat16* xp = X + (static_cast<size_t>(b) * HW + p0 + q) * ldx + c0 + v * 8;
    int n = p1 - p0 - q;
    n = n > 0 ? (n + ppar - 1) / ppar : 0;
    auto acc8 = [&](const uint4& u) {
      const float2 a = unpack_bf16(u.x), b2 = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      const float f[8] = {a.x, a.y, b2.x, b2.y, c.x, c.y, d.x, d.y};
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; ss[i] = fmaf(f[i], f[i], ss[i]); }
    };
    for (; n >= U; n -= U) {            // U independent 16-byte loads in flight, no bounds predicates
      uint4 u[U];
#pragma unroll
      for (int k = 0; k < U; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(xp + k * xstep));
#pragma unroll
      for (int k = 0; k < U; ++k) acc8(u[k]);
      xp += U * xstep;
    }
    if (n > 0) {                        // remainder: one more batch of (predicated) loads, not n round trips
      uint4 u[U];
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (k < n) u[k] = __ldg(reinterpret_cast<const uint4*>(xp + k * xstep));
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (k < n) acc8(u[k]);
    }
    float* row = sh + static_cast<size_t>(q) * 2 * cslab;
#pragma unroll
    for (int i = 0; i < 8; ++i) { row[v * 8 + i] = s[i]; row[cslab + v * 8 + i] = ss[i]; }
  }
  __syncthreads();
  // fold rows then channels into groups; one thread per group that intersects this slab
  const int cpg = C / groups;
  const int g_first = c0 / cpg;
  const int g_last = (min(C, c0 + cslab) - 1) / cpg;
  for (int g = g_first + threadIdx.x; g <= g_last; g += blockDim.x) {
    const int lo = max(g * cpg, c0) - c0, hi = min((g + 1) * cpg, min(C, c0 + cslab)) - c0;
    float a = 0.f, a2 = 0.f;
    for (int r = 0; r < ppar; ++r) {
      const float* row = sh + static_cast<size_t>(r) * 2 * cslab;
      for (int c = lo; c < hi; ++c) { a += row[c]; a2 += row[cslab + c]; }
    }
    float* dst = partial + ((((static_cast<size_t>(b) * gridDim.x + blockIdx.x) * gridDim.z + blockIdx.z) * groups + g) * 2);
    dst[0] = a;
    dst[1] = a2;
  }
}

// ------------------------------------------------------------ GroupNorm apply
// Same thread <-> channel-vector mapping, so the 8 scale/shift pairs of a thread live in
// registers.  Every CTA first re-reduces the (tiny) partial-sum table in a fixed order.
template <int U>
__global__ void __launch_bounds__(256)
groupnorm_apply_kernel(const __nv_bfloat16* __restrict__ X, int ldx, __nv_bfloat16* __restrict__ Y, int ldy,
                       int HW, int C, int groups, int rows_per_cta, int vslab, int nchunks_stats, int nslab_stats,
                       int cslab_stats, const float* __restrict__ partial, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float eps, int silu, int direct) {
  __shared__ float g_mean[64], g_rstd[64];  // groups intersecting this slab (<= 64)
  const int cslab = vslab * 8;
  const int ppar = blockDim.x / vslab;
  const int v = threadIdx.x % vslab;
  const int q = threadIdx.x / vslab;
  const int b = blockIdx.y;
  const int c0 = blockIdx.z * cslab;
  const int cpg = C / groups;
  const int g_first = c0 / cpg;
  const int g_last = (min(C, c0 + cslab) - 1) / cpg;
  const float inv_n = 1.f / (static_cast<float>(cpg) * static_cast<float>(HW));
  // The affine parameters do not depend on the previous kernel: fetch them before the programmatic-dependent-launch
  // wait, so their latency is hidden under the predecessor's tail.
  const int cbase = c0 + v * 8;
  float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga, ba = ga, bb = ga;
  if (q < ppar) {
    ga = __ldg(reinterpret_cast<const float4*>(gamma + cbase));
    gb = __ldg(reinterpret_cast<const float4*>(gamma + cbase + 4));
    ba = __ldg(reinterpret_cast<const float4*>(beta + cbase));
    bb = __ldg(reinterpret_cast<const float4*>(beta + cbase + 4));
  }
  pdl_wait();
  if (direct) {
    // `partial` holds given statistics [B][groups][2] = (mean, biased variance): tiled VAE, pooled over tiles
    for (int g = g_first + threadIdx.x; g <= g_last; g += blockDim.x) {
      const float2 mv = __ldg(reinterpret_cast<const float2*>(partial + (static_cast<size_t>(b) * groups + g) * 2));
      g_mean[g - g_first] = mv.x;
      g_rstd[g - g_first] = rsqrtf(mv.y + eps);
    }
  } else {
    // Four lanes per group stride over that group's (chunk, slab) partial sums with four loads in flight each, then a
    // fixed two-step shuffle tree: every group of the slab is reduced at the same time, so the whole table costs about
    // one memory round trip.  Only full warps take part (the block size need not be a multiple of 32).
    const int nfull = static_cast<int>(blockDim.x) & ~31;
    const int ng = g_last - g_first + 1;
    const int sub = threadIdx.x & 3;
    for (int gbase = 0; gbase < ng; gbase += nfull >> 2) {
      const int gl = gbase + (static_cast<int>(threadIdx.x) >> 2);
      if (static_cast<int>(threadIdx.x) < nfull) {
        float a = 0.f, a2 = 0.f;
        if (gl < ng) {
          const int g = g_first + gl;
          const int s_lo = (g * cpg) / cslab_stats, s_hi = ((g + 1) * cpg - 1) / cslab_stats;
          const int ns = s_hi - s_lo + 1;
          const int n = nchunks_stats * ns;
          for (int i = sub; i < n; i += 16) {
            float2 pr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int idx = i + 4 * k;
              pr[k] = make_float2(0.f, 0.f);
              if (idx < n) {
                const int ch = idx / ns, sl = s_lo + (idx - ch * ns);
                pr[k] = __ldg(reinterpret_cast<const float2*>(
                    partial + ((((static_cast<size_t>(b) * nchunks_stats + ch) * nslab_stats + sl) * groups + g) * 2)));
              }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { a += pr[k].x; a2 += pr[k].y; }
          }
        }
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
        if (sub == 0 && gl < ng) {
          const float mean = a * inv_n;
          const float var = fmaxf(a2 * inv_n - mean * mean, 0.f);
          g_mean[gl] = mean;
          g_rstd[gl] = rsqrtf(var + eps);
        }
      }
    }
  }
  __syncthreads();
  if (q < ppar) {
    float sc[8], sf[8];
    {
      const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
      const float bt[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int g = (cbase + i) / cpg - g_first;
        sc[i] = g_rstd[g] * gg[i];
        sf[i] = bt[i] - g_mean[g] * sc[i];
      }
    }
    const int p0 = blockIdx.x * rows_per_cta;
    const int p1 = min(HW, p0 + rows_per_cta);
    // pointer-stepped loop (no per-vector index arithmetic or bounds predicates in the main body): this thread owns
    // vectors p0 + q, p0 + q + ppar, ... of its channel column
    const size_t xstep = static_cast<size_t>(ppar) * ldx, ystep = static_cast<size_t>(ppar) * ldy;
    const __nv_bfloat16* xp = X + (static_cast<size_t>(b) * HW + p0 + q) * ldx + c0 + v * 8;
    __nv_bfloat16* yp = Y + (static_cast<size_t>(b) * HW + p0 + q) * ldy + c0 + v * 8;
    int n = p1 - p0 - q;
    n = n > 0 ? (n + ppar - 1) / ppar : 0;
    const bool do_silu = silu != 0;
    for (; n >= U; n -= U) {
      uint4 u[U];
#pragma unroll
      for (int k = 0; k < U; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(xp + k * xstep));
#pragma unroll
      for (int k = 0; k < U; ++k) *reinterpret_cast<uint4*>(yp + k * ystep) = norm8(u[k], sc, sf, do_silu);
      xp += U * xstep;
      yp += U * ystep;
    }
    if (n > 0) {                        // remainder: one more batch of (predicated) loads, not n round trips
      uint4 u[U];
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (k < n) u[k] = __ldg(reinterpret_cast<const uint4*>(xp + k * xstep));
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (k < n) *reinterpret_cast<uint4*>(yp + k * ystep) = norm8(u[k], sc, sf, do_silu);
    }
  }
  pdl_launch_dependents();
}

// ------------------------------------------------- GroupNorm statistics pooling (tiled VAE)
// Folds the partial sums of one tile (groupnorm_stats_kernel) into that tile's (mean, biased variance) per
// (image, group) and accumulates weight * (mean, var) into acc[B][groups][2] — the pixel-weighted average of
// per-tile means AND per-tile variances of GroupNormParam.summary (utils/tilevae/tilevae.py:263-278).  One
// thread per (image, group), fixed order, no atomics: launches on one stream accumulate deterministically.
__global__ void groupnorm_pool_kernel(const float* __restrict__ partial, int B, int HW, int C, int groups,
                                      int nchunks, int nslab, int cslab, float weight, float* __restrict__ acc) {
  PdlScope pdl_scope;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * groups) return;
  const int b = i / groups, g = i - b * groups;
  const int cpg = C / groups;
  const int s_lo = (g * cpg) / cslab, s_hi = ((g + 1) * cpg - 1) / cslab;
  float a = 0.f, a2 = 0.f;
  for (int ch = 0; ch < nchunks; ++ch)
    for (int sl = s_lo; sl <= s_hi; ++sl) {
      const float2 pr = __ldg(reinterpret_cast<const float2*>(
          partial + ((((static_cast<size_t>(b) * nchunks + ch) * nslab + sl) * groups + g) * 2)));
      a += pr.x;
      a2 += pr.y;
    }
  const float inv_n = 1.f / (static_cast<float>(cpg) * static_cast<float>(HW));
  const float mean = a * inv_n;
  const float var = fmaxf(a2 * inv_n - mean * mean, 0.f);
  acc[2 * i] += weight * mean;
  acc[2 * i + 1] += weight * var;
}

// ------------------------------------------------- GroupNorm statistics from GEMM-epilogue partial sums
// partial = [B][slabs][C/4][2] (sum, sum of squares) per (32-row slab, 4-channel unit), written by the epilogue of the
// convolution / GEMM that produced the tensor (gemm2.cu: gn_partial_store).  One CTA per (image, four consecutive units =
// 32 contiguous bytes per slab): thread t sums slabs t, t + 512, ... (eight 2 x 16-byte loads in flight), then a fixed
// shuffle tree + a fixed-order sum over the 16 warps; cpg / 4 units form a group (cpg in {4, 8, 16}, so a group never
// leaves the CTA).  Writes (mean, biased variance) per (image, group): the input of groupnorm_apply_kernel's direct mode.
__global__ void __launch_bounds__(512)
groupnorm_fold_kernel(const float* __restrict__ partial, int slabs, int units, int groups, int cpg, float inv_n,
                      float* __restrict__ mean_var) {
  PdlScope pdl_scope;
  __shared__ float red[16][8];
  const int b = blockIdx.y, u0 = blockIdx.x * 4;
  const float* src = partial + (static_cast<size_t>(b) * slabs * units + u0) * 2;
  const size_t sstride = static_cast<size_t>(units) * 2;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int s0 = threadIdx.x; s0 < slabs; s0 += 8 * 512) {
    float4 lo[8], hi[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int s = s0 + k * 512;
      lo[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      hi[k] = lo[k];
      if (s < slabs) {
        const float4* q = reinterpret_cast<const float4*>(src + static_cast<size_t>(s) * sstride);
        lo[k] = __ldg(q);
        hi[k] = __ldg(q + 1);
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[0] += lo[k].x; acc[1] += lo[k].y; acc[2] += lo[k].z; acc[3] += lo[k].w;
      acc[4] += hi[k].x; acc[5] += hi[k].y; acc[6] += hi[k].z; acc[7] += hi[k].w;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = warp_sum(acc[i]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][i] = acc[i];
  }
  __syncthreads();
  const int upg = cpg >> 2;            // units per group: 1, 2 or 4
  const int ng = 4 / upg;              // groups of this CTA
  if (static_cast<int>(threadIdx.x) < ng) {
    float s1 = 0.f, s2 = 0.f;
    for (int w = 0; w < 16; ++w)
      for (int u = 0; u < upg; ++u) {
        s1 += red[w][2 * (threadIdx.x * upg + u)];
        s2 += red[w][2 * (threadIdx.x * upg + u) + 1];
      }
    const int g = u0 / upg + threadIdx.x;
    const float mean = s1 * inv_n;
    const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);
    reinterpret_cast<float2*>(mean_var)[static_cast<size_t>(b) * groups + g] = make_float2(mean, var);
  }
}

// General form (any group width that is a multiple of the unit, few slabs: the UNet / ControlNet tensors): one CTA of
// four warps per (image, group); thread t sums the group's (slab, unit) pairs t, t + 128, ... (a group's units are
// contiguous: upg * 8 bytes per slab), fixed shuffle tree + fixed-order sum over the warps.
__global__ void __launch_bounds__(128)
groupnorm_fold_general_kernel(const float* __restrict__ partial, int slabs, int units, int groups, int upg, float inv_n,
                              float* __restrict__ mean_var) {
  PdlScope pdl_scope;
  __shared__ float red[4][2];
  const int g = blockIdx.x, b = blockIdx.y;
  const float2* src = reinterpret_cast<const float2*>(partial) + static_cast<size_t>(b) * slabs * units + g * upg;
  const int n = slabs * upg;
  float s1 = 0.f, s2 = 0.f;
  for (int i0 = threadIdx.x; i0 < n; i0 += 4 * 128) {
    float2 t[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + k * 128;
      t[k] = make_float2(0.f, 0.f);
      if (i < n) {
        const int sl = i / upg, u = i - sl * upg;
        t[k] = __ldg(src + static_cast<size_t>(sl) * units + u);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { s1 += t[k].x; s2 += t[k].y; }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[warp][0] = s1; red[warp][1] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float a = (red[0][0] + red[1][0]) + (red[2][0] + red[3][0]);
    const float a2 = (red[0][1] + red[1][1]) + (red[2][1] + red[3][1]);
    const float mean = a * inv_n;
    const float var = fmaxf(a2 * inv_n - mean * mean, 0.f);
    reinterpret_cast<float2*>(mean_var)[static_cast<size_t>(b) * groups + g] = make_float2(mean, var);
  }
}

// ------------------------------------------------------------------ LayerNorm
// One warp per row, the row lives in registers (two exact passes).
template <int MAXV>
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ X, int ldx, __nv_bfloat16* __restrict__ Y,
                                 int ldy, int M, int C, int c_real, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps) {
  // c_real <= C: the statistics run over the first c_real channels only; channels [c_real, C) are padding that holds
  // zeros on input and (gamma = beta = 0 there) on output — SwinIR's 180 channels live in 192-wide rows.
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int vpr = C / 8;
  float f[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int v = lane + 32 * k;
    if (v < vpr) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(X + static_cast<size_t>(row) * ldx + v * 8));
      unpack8(u, f[k]);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += f[k][i];
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(c_real);
  float s2 = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int v = lane + 32 * k;
    if (v < vpr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = (v * 8 + i < c_real) ? f[k][i] - mean : 0.f;
        s2 += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(s2) / static_cast<float>(c_real) + eps);
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int v = lane + 32 * k;
    if (v < vpr) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + v * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (f[k][i] - mean) * rstd * gg[i] + bb[i];
      *reinterpret_cast<uint4*>(Y + static_cast<size_t>(row) * ldy + v * 8) = pack8(o);
    }
  }
}

// ---------------------------------------------------------------- row softmax
// One CTA per row; fp32 logits in, bf16 probabilities out.
__global__ void softmax_rows_kernel(const float* __restrict__ S, int lds, __nv_bfloat16* __restrict__ P,
                                    int ldp, int N, float scale_log2) {
  PdlScope pdl_scope;  // PDL: wait for the previous kernel on entry, trigger the next one on exit
  __shared__ float red[32];
  const int row = blockIdx.x;
  const float* s = S + static_cast<size_t>(row) * lds;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < N; i += blockDim.x) mx = fmaxf(mx, s[i]);
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = lane < nw ? red[lane] : -INFINITY;
  mx = warp_max(mx);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) sum += exp2f((s[i] - mx) * scale_log2);
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = lane < nw ? red[lane] : 0.f;
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  __nv_bfloat16* p = P + static_cast<size_t>(row) * ldp;
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    p[i] = __float2bfloat16(exp2f((s[i] - mx) * scale_log2) * inv);
}


// ------------------------------------------------- GroupNorm, single launch (cluster per image)
// For tensors that stay in L2 (every GroupNorm of the UNet / ControlNet): a cluster of `CS` CTAs owns one
// image.  Pass 1: each CTA sums its pixel range (thread <-> channel-vector mapping as above, eight 16-byte
// loads in flight), folds them to per-group partials in shared memory; one cluster barrier; every CTA reads
// the CS partial tables through distributed shared memory in rank order (deterministic), and pass 2 normalises
// (+SiLU) the rows it kept in registers (or re-reads them: L2 hits) and stores.  Replaces two launches and the partial-sum round trip.
__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t cluster_addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
  return v;
}

__global__ void __launch_bounds__(512)
groupnorm_fused_kernel(const __nv_bfloat16* __restrict__ X, int ldx, __nv_bfloat16* __restrict__ Y, int ldy,
                       int HW, int C, int groups, int cs, const float* __restrict__ gamma,
                       const float* __restrict__ beta, float eps, int silu, int smem_rows) {
  extern __shared__ __align__(16) float sh[];   // [ppar][2][C] per-thread-row channel sums, then (smem_rows > 0) the CTA's
                                                // slice of the image, [smem_rows][C] bf16, kept between the two passes
  __shared__ float part[128];              // this CTA's per-group (sum, sumsq)
  __shared__ float g_mean[64], g_rstd[64];
  const int vpr = C >> 3;
  const int ppar = blockDim.x / vpr;
  const int v = threadIdx.x % vpr;
  const int q = threadIdx.x / vpr;
  const int rank = static_cast<int>(cluster_ctarank());
  const int b = blockIdx.y;
  const int rows = (HW + cs - 1) / cs;
  const int p0 = rank * rows, p1 = min(HW, p0 + rows);
  const __nv_bfloat16* xb = X + (static_cast<size_t>(b) * HW) * ldx + v * 8;
  // the affine parameters do not depend on the previous kernel: fetched before the PDL wait
  const int cbase = v * 8;
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + cbase));
  const float4 gb = __ldg(reinterpret_cast<const float4*>(gamma + cbase + 4));
  const float4 ba = __ldg(reinterpret_cast<const float4*>(beta + cbase));
  const float4 bb = __ldg(reinterpret_cast<const float4*>(beta + cbase + 4));
  pdl_wait();
  // Up to 16 vectors per thread (every shape the launch geometry admits today) stay in registers between the two
  // passes, so the tensor is read once; larger slices fall back to re-reading their rows (L2 hits) in pass 2.
  constexpr int kIt = 2;
  const bool in_regs = rows <= kIt * 8 * ppar;
  // Mid-size tensors (1 - 2.75 MB per image: the 64x64 and 32x32 levels): the slice does not fit the registers but it
  // fits shared memory (<= 168 KB of the 227 KB), so it is still read from L2 exactly once.
  const bool in_smem = !in_regs && smem_rows >= p1 - p0;
  uint4* sdata = reinterpret_cast<uint4*>(sh + static_cast<size_t>(ppar) * 2 * C);
  uint4 keep[kIt][8];
  {
    float s[8], ss[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = ss[i] = 0.f;
    if (in_regs) {
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int pidx = p0 + q + it * 8 * ppar;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          keep[it][k] = make_uint4(0u, 0u, 0u, 0u);   // bf16 zeros add nothing to either sum
          if (pidx + k * ppar < p1)
            keep[it][k] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(pidx + k * ppar) * ldx));
        }
      }
#pragma unroll
      for (int it = 0; it < kIt; ++it)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float f[8];
          unpack8(keep[it][k], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) { s[i] += f[i]; ss[i] += f[i] * f[i]; }
        }
    } else {
      for (int pidx = p0 + q; pidx < p1; pidx += 8 * ppar) {
        uint4 u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          u[k] = make_uint4(0u, 0u, 0u, 0u);
          if (pidx + k * ppar < p1)
            u[k] = __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(pidx + k * ppar) * ldx));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float f[8];
          unpack8(u[k], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) { s[i] += f[i]; ss[i] += f[i] * f[i]; }
          if (in_smem && pidx + k * ppar < p1) sdata[static_cast<size_t>(pidx + k * ppar - p0) * vpr + v] = u[k];
        }
      }
    }
    float* row = sh + static_cast<size_t>(q) * 2 * C;
#pragma unroll
    for (int i = 0; i < 8; ++i) { row[v * 8 + i] = s[i]; row[C + v * 8 + i] = ss[i]; }
  }
  __syncthreads();
  const int cpg = C / groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int g = warp; g < groups; g += nwarps) {   // one warp per group, fixed shuffle tree
    float a = 0.f, a2 = 0.f;
    for (int i = lane; i < ppar * cpg; i += 32) {
      const int r = i / cpg, c = g * cpg + (i - r * cpg);
      const float* row = sh + static_cast<size_t>(r) * 2 * C;
      a += row[c];
      a2 += row[C + c];
    }
    a = warp_sum(a);
    a2 = warp_sum(a2);
    if (lane == 0) { part[2 * g] = a; part[2 * g + 1] = a2; }
  }
  __syncthreads();
  cluster_arrive_release();
  cluster_wait_acquire();
  if (static_cast<int>(threadIdx.x) < groups) {
    const uint32_t local = smem_u32(&part[2 * threadIdx.x]);
    float a = 0.f, a2 = 0.f;
    for (int r = 0; r < cs; ++r) {
      const uint32_t remote = mapa_u32(local, static_cast<uint32_t>(r));
      a += ld_dsmem_f32(remote);
      a2 += ld_dsmem_f32(remote + 4);
    }
    const float inv_n = 1.f / (static_cast<float>(cpg) * static_cast<float>(HW));
    const float mean = a * inv_n;
    const float var = fmaxf(a2 * inv_n - mean * mean, 0.f);
    g_mean[threadIdx.x] = mean;
    g_rstd[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  cluster_arrive_release();   // this CTA is done reading its peers' tables (waited on before exit)
  float sc[8], sf[8];
  {
    const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    const float bt[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (cbase + i) / cpg;
      sc[i] = g_rstd[g] * gg[i];
      sf[i] = bt[i] - g_mean[g] * sc[i];
    }
  }
  __nv_bfloat16* yb = Y + (static_cast<size_t>(b) * HW) * ldy + v * 8;
  const bool do_silu = silu != 0;
  if (in_regs) {
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int pidx = p0 + q + it * 8 * ppar;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (pidx + k * ppar < p1)
          *reinterpret_cast<uint4*>(yb + static_cast<size_t>(pidx + k * ppar) * ldy) = norm8(keep[it][k], sc, sf, do_silu);
      }
    }
  } else {
    for (int pidx = p0 + q; pidx < p1; pidx += 8 * ppar) {
      uint4 u[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (pidx + k * ppar < p1)
          u[k] = in_smem ? sdata[static_cast<size_t>(pidx + k * ppar - p0) * vpr + v]   // written by this same thread
                         : __ldg(reinterpret_cast<const uint4*>(xb + static_cast<size_t>(pidx + k * ppar) * ldx));
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (pidx + k * ppar < p1)
          *reinterpret_cast<uint4*>(yb + static_cast<size_t>(pidx + k * ppar) * ldy) = norm8(u[k], sc, sf, do_silu);
      }
    }
  }
  cluster_wait_acquire();     // no CTA may exit while a peer can still read its shared memory
  pdl_launch_dependents();
}

// Launch geometry of the fused kernel; cs == 0: not eligible (use the two-pass kernels).
struct GnFusedGeom {
  int cs, threads;
  size_t smem;
  int smem_rows;   // rows of the image slice kept in shared memory between the passes (0: registers / re-read)
};
constexpr size_t kGnFusedSmemMax = 220 * 1024;
static GnFusedGeom gn_fused_geom(int B, int HW, int C, int groups) {
  GnFusedGeom g{0, 0, 0, 0};
  if (C % 8 != 0 || groups <= 0 || groups > 64 || C % groups != 0) return g;
  const int vpr = C / 8;
  if (vpr > 512) return g;
  const size_t bytes_per_image = static_cast<size_t>(HW) * C * 2;
  // Small L2-resident tensors only: measured on B200 the single launch wins up to ~1 MB per image (8x8 / 16x16
  // levels); above that the 8 CTAs per image cannot pull enough bandwidth and the two-pass kernels are faster.
  // (16-CTA non-portable clusters for the 1 - 2.75 MB tensors were measured too: no gain, profiles/r01g.)
  static const bool smem_path = [] {
    // Opt-in experiment (EDTR_GN_SMEM=1): 16-CTA clusters with the image slice resident in shared memory for the
    // 1 - 2.75 MB tensors.  Measured on B200 (profiles/r02p_groupnorm_smem_cluster.txt): 41 us vs 33 us for the two
    // launches of the two-pass kernels at 8 x 64 x 64 x 320 — eight 16-CTA clusters do not fill the GPU.  Off.
    const char* e = getenv("EDTR_GN_SMEM");
    return e != nullptr && e[0] == '1';
  }();
  // above 1 MB per image the slice of a CTA no longer fits the registers: a 16-CTA (non-portable) cluster per image
  // keeps it in shared memory instead, up to 2.75 MB per image (16 x 168 KB)
  const size_t limit = smem_path ? (2816u << 10) : (1u << 20);
  if (bytes_per_image > limit || static_cast<size_t>(B) * bytes_per_image > (64u << 20)) return g;
  const int ppar = 512 / vpr;
  g.threads = vpr * ppar;
  if (g.threads % 32 != 0) {            // whole warps only (warp-per-group reduction, .aligned barriers)
    int t = g.threads / 32 * 32;
    if (t < vpr) return g;
    g.threads = t / vpr * vpr;
    if (g.threads % 32 != 0) return g;
  }
  g.smem = static_cast<size_t>(g.threads / vpr) * 2 * C * sizeof(float);
  g.cs = HW >= 64 ? 8 : (HW >= 8 ? 2 : 1);
  if (bytes_per_image > (1u << 20)) {
    const int pp = g.threads / vpr;
    for (int cs = 8; cs <= 16; cs *= 2) {
      const int rows = (HW + cs - 1) / cs;
      if (rows <= 2 * 8 * pp) { g.cs = cs; break; }                       // fits the registers after all
      const size_t need = g.smem + static_cast<size_t>(rows) * C * 2;
      if (need <= kGnFusedSmemMax) { g.cs = cs; g.smem = need; g.smem_rows = rows; break; }
      if (cs == 16) { g.cs = 0; return g; }
    }
  }
  return g;
}

struct GnGeom {
  int vslab, nslab, ppar, threads, rows, nchunks;
  bool stream;   // HBM-streaming tensor (>= 64 MB): 8 CTAs per SM, four loads in flight per thread
};

// Launch geometry shared by the stats and apply kernels (and by edtr_groupnorm_partial_size).
static GnGeom gn_geom(int B, int HW, int C) {
  GnGeom g;
  const int vpr = C / 8;
  g.vslab = vpr;
  if (g.vslab > 256) {
    g.vslab = 256;
    while (vpr % g.vslab != 0) --g.vslab;
  }
  g.nslab = vpr / g.vslab;
  g.ppar = 256 / g.vslab;
  if (g.ppar < 1) g.ppar = 1;
  g.threads = g.vslab * g.ppar;
  // tensors that stream from HBM (VAE decoder, >= 64 MB) get 8 CTAs per SM so that enough 16-byte loads are in
  // flight to cover the DRAM latency
  const size_t bytes = static_cast<size_t>(B) * HW * C * 2;
  g.stream = bytes >= (64u << 20);
  int rows;
  if (g.stream) {
    const int want_ctas = 8 * 148;
    const int chunks = (want_ctas + B * g.nslab - 1) / (B * g.nslab);
    rows = (HW + chunks - 1) / chunks;
    if (rows < 4 * g.ppar) rows = 4 * g.ppar;
  } else {
    // L2-resident tensors: the launches are bound by latency, not bandwidth.  One pass of eight 16-byte loads per
    // thread when that needs no more than 64 chunks per image (the apply kernel re-reduces chunks x groups partial
    // sums per CTA), else as few passes as 64 chunks allow.
    rows = 8 * g.ppar;
    const int cap_rows = (HW + 63) / 64;
    if (rows < cap_rows) rows = cap_rows;
  }
  if (rows > HW) rows = HW;
  g.rows = rows;
  g.nchunks = (HW + rows - 1) / rows;
  return g;
}

int prime_norm_attributes() {
  cudaError_t e = cudaFuncSetAttribute(groupnorm_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(kGnFusedSmemMax));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(groupnorm_fused_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(groupnorm_fused): %s", cudaGetErrorString(e));
    return EDTR_ERR_CUDA;
  }
  return EDTR_OK;
}

}  // namespace edtr

using namespace edtr;

static int check_gn_args(const void* X, int ldx, int B, int HW, int C, int groups) {
  EDTR_REQUIRE(X != nullptr, "X is NULL");
  EDTR_REQUIRE(B > 0 && HW > 0 && C > 0 && groups > 0, "bad GroupNorm shape");
  EDTR_REQUIRE(C % 8 == 0 && C % groups == 0, "C (%d) must be a multiple of 8 and of groups (%d)", C, groups);
  EDTR_REQUIRE(ldx % 8 == 0 && ldx >= C, "ldx must be >= C and a multiple of 8");
  EDTR_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0, "X must be 16-byte aligned");
  EDTR_REQUIRE(B <= 65535, "B too large");
  return EDTR_OK;
}

extern "C" size_t edtr_groupnorm_partial_size(int B, int HW, int C, int groups) {
  if (B <= 0 || HW <= 0 || C <= 0 || groups <= 0 || C % 8 != 0) return 0;
  const GnGeom g = gn_geom(B, HW, C);
  return static_cast<size_t>(B) * g.nchunks * g.nslab * groups * 2;
}

extern "C" int edtr_groupnorm_stats(const void* X, int ldx, int B, int HW, int C, int groups, float* stats,
                                    void* stream) {
  int rc = check_gn_args(X, ldx, B, HW, C, groups);
  if (rc) return rc;
  EDTR_REQUIRE(stats != nullptr, "stats is NULL");
  const GnGeom g = gn_geom(B, HW, C);
  dim3 grid(g.nchunks, B, g.nslab);
  const size_t sh = static_cast<size_t>(g.ppar) * 2 * g.vslab * 8 * sizeof(float);
  EDTR_REQUIRE(sh <= 48 * 1024, "GroupNorm stats shared memory too large");
  if (g.stream)
    EDTR_LAUNCH(groupnorm_stats_kernel<4>, grid, g.threads, sh, static_cast<cudaStream_t>(stream),
                reinterpret_cast<const __nv_bfloat16*>(X), ldx, HW, C, groups, g.rows, g.vslab, stats);
  else
    EDTR_LAUNCH(groupnorm_stats_kernel<8>, grid, g.threads, sh, static_cast<cudaStream_t>(stream),
                reinterpret_cast<const __nv_bfloat16*>(X), ldx, HW, C, groups, g.rows, g.vslab, stats);
  return check_launch("groupnorm_stats_kernel");
}

extern "C" int edtr_groupnorm_apply(const void* X, int ldx, void* Y, int ldy, int B, int HW, int C, int groups,
                                    const float* stats, const float* gamma, const float* beta, float eps,
                                    int silu, void* stream) {
  int rc = check_gn_args(X, ldx, B, HW, C, groups);
  if (rc) return rc;
  EDTR_REQUIRE(Y && stats && gamma && beta, "Y/stats/gamma/beta is NULL");
  EDTR_REQUIRE(ldy % 8 == 0 && ldy >= C && (reinterpret_cast<uintptr_t>(Y) & 15) == 0, "bad Y stride/alignment");
  EDTR_REQUIRE(((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
               "gamma/beta must be 16-byte aligned");
  const GnGeom g = gn_geom(B, HW, C);
  EDTR_REQUIRE(g.vslab * 8 / (C / groups) + 2 <= 64, "too many groups per channel slab");
  dim3 grid(g.nchunks, B, g.nslab);
  if (g.stream)
    EDTR_LAUNCH(groupnorm_apply_kernel<4>, grid, g.threads, 0, static_cast<cudaStream_t>(stream),
                reinterpret_cast<const __nv_bfloat16*>(X), ldx, reinterpret_cast<__nv_bfloat16*>(Y), ldy, HW, C, groups,
                g.rows, g.vslab, g.nchunks, g.nslab, g.vslab * 8, stats, gamma, beta, eps, silu, 0);
  else
    EDTR_LAUNCH(groupnorm_apply_kernel<8>, grid, g.threads, 0, static_cast<cudaStream_t>(stream),
                reinterpret_cast<const __nv_bfloat16*>(X), ldx, reinterpret_cast<__nv_bfloat16*>(Y), ldy, HW, C, groups,
                g.rows, g.vslab, g.nchunks, g.nslab, g.vslab * 8, stats, gamma, beta, eps, silu, 0);
  return check_launch("groupnorm_apply_kernel");
}

extern "C" int edtr_groupnorm_apply_stats(const void* X, int ldx, void* Y, int ldy, int B, int HW, int C, int groups,
                                          const float* mean_var, const float* gamma, const float* beta, float eps,
                                          int silu, void* stream) {
  int rc = check_gn_args(X, ldx, B, HW, C, groups);
  if (rc) return rc;
  EDTR_REQUIRE(Y && mean_var && gamma && beta, "Y/mean_var/gamma/beta is NULL");
  EDTR_REQUIRE(ldy % 8 == 0 && ldy >= C && (reinterpret_cast<uintptr_t>(Y) & 15) == 0, "bad Y stride/alignment");
  EDTR_REQUIRE(((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
               "gamma/beta must be 16-byte aligned");
  EDTR_REQUIRE((reinterpret_cast<uintptr_t>(mean_var) & 7) == 0, "mean_var must be 8-byte aligned");
  const GnGeom g = gn_geom(B, HW, C);
  EDTR_REQUIRE(g.vslab * 8 / (C / groups) + 2 <= 64, "too many groups per channel slab");
  dim3 grid(g.nchunks, B, g.nslab);
  if (g.stream)
    EDTR_LAUNCH(groupnorm_apply_kernel<4>, grid, g.threads, 0, static_cast<cudaStream_t>(stream),
                reinterpret_cast<const __nv_bfloat16*>(X), ldx, reinterpret_cast<__nv_bfloat16*>(Y), ldy, HW, C, groups,
                g.rows, g.vslab, g.nchunks, g.nslab, g.vslab * 8, mean_var, gamma, beta, eps, silu, 1);
  else
    EDTR_LAUNCH(groupnorm_apply_kernel<8>, grid, g.threads, 0, static_cast<cudaStream_t>(stream),
                reinterpret_cast<const __nv_bfloat16*>(X), ldx, reinterpret_cast<__nv_bfloat16*>(Y), ldy, HW, C, groups,
                g.rows, g.vslab, g.nchunks, g.nslab, g.vslab * 8, mean_var, gamma, beta, eps, silu, 1);
  return check_launch("groupnorm_apply_kernel");
}

extern "C" int edtr_groupnorm_pool(const float* stats, int B, int HW, int C, int groups, float weight, float* acc,
                                   void* stream) {
  EDTR_REQUIRE(stats && acc, "stats/acc is NULL");
  EDTR_REQUIRE(B > 0 && HW > 0 && C > 0 && groups > 0 && C % 8 == 0 && C % groups == 0, "bad GroupNorm shape");
  const GnGeom g = gn_geom(B, HW, C);
  const int n = B * groups;
  EDTR_LAUNCH(groupnorm_pool_kernel, (n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), stats, B, HW, C,
              groups, g.nchunks, g.nslab, g.vslab * 8, weight, acc);
  return check_launch("groupnorm_pool_kernel");
}

extern "C" int edtr_groupnorm_fold(const float* gn_partial, int B, int slabs, int C, int groups, int unit,
                                   float* mean_var, void* stream) {
  EDTR_REQUIRE(gn_partial && mean_var, "gn_partial/mean_var is NULL");
  EDTR_REQUIRE(B > 0 && B <= 65535 && slabs > 0 && C > 0 && groups > 0 && groups <= 65535 && C % groups == 0,
               "bad GroupNorm shape");
  EDTR_REQUIRE(unit == 2 || unit == 4, "unit must be 2 or 4 (got %d)", unit);
  const int cpg = C / groups;
  EDTR_REQUIRE(cpg % unit == 0, "edtr_groupnorm_fold needs (C / groups) %% unit == 0 (C/groups %d, unit %d)", cpg, unit);
  EDTR_REQUIRE(((reinterpret_cast<uintptr_t>(gn_partial) & 31) == 0) && ((reinterpret_cast<uintptr_t>(mean_var) & 7) == 0),
               "gn_partial must be 32-byte aligned, mean_var 8-byte aligned");
  const int units = C / unit;
  const float inv_n = 1.f / (static_cast<float>(cpg) * 32.f * static_cast<float>(slabs));
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (unit == 4 && (cpg == 4 || cpg == 8 || cpg == 16) && slabs >= 256) {
    // the streaming VAE tensors: four-unit blocks (C % 16 == 0 here), 32 contiguous bytes per slab and thread
    EDTR_LAUNCH(groupnorm_fold_kernel, dim3(units / 4, B), 512, 0, st, gn_partial, slabs, units, groups, cpg, inv_n, mean_var);
    return check_launch("groupnorm_fold_kernel");
  }
  EDTR_LAUNCH(groupnorm_fold_general_kernel, dim3(groups, B), 128, 0, st, gn_partial, slabs, units, groups, cpg / unit,
              inv_n, mean_var);
  return check_launch("groupnorm_fold_general_kernel");
}

extern "C" int edtr_groupnorm_fused_supported(int B, int HW, int C, int groups) {
  if (B <= 0 || HW <= 0 || C <= 0 || B > 65535) return 0;
  return gn_fused_geom(B, HW, C, groups).cs > 0 ? 1 : 0;
}

extern "C" int edtr_groupnorm_fused(const void* X, int ldx, void* Y, int ldy, int B, int HW, int C, int groups,
                                    const float* gamma, const float* beta, float eps, int silu, void* stream) {
  int rc = check_gn_args(X, ldx, B, HW, C, groups);
  if (rc) return rc;
  EDTR_REQUIRE(Y && gamma && beta, "Y/gamma/beta is NULL");
  EDTR_REQUIRE(ldy % 8 == 0 && ldy >= C && (reinterpret_cast<uintptr_t>(Y) & 15) == 0, "bad Y stride/alignment");
  EDTR_REQUIRE(((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
               "gamma/beta must be 16-byte aligned");
  const GnFusedGeom g = gn_fused_geom(B, HW, C, groups);
  EDTR_REQUIRE(g.cs > 0, "shape not supported by the fused GroupNorm (see edtr_groupnorm_fused_supported)");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(g.cs, B, 1);
  cfg.blockDim = dim3(g.threads, 1, 1);
  cfg.dynamicSmemBytes = g.smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = g.cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  (void)cudaLaunchKernelEx(&cfg, groupnorm_fused_kernel, reinterpret_cast<const __nv_bfloat16*>(X), ldx,
                           reinterpret_cast<__nv_bfloat16*>(Y), ldy, HW, C, groups, g.cs, gamma, beta, eps, silu,
                           g.smem_rows);
  return check_launch("groupnorm_fused_kernel");
}

extern "C" int edtr_layernorm_bf16(const void* X, int ldx, void* Y, int ldy, int M, int C, const float* gamma,
                                   const float* beta, float eps, void* stream) {
  return edtr_layernorm_padded_bf16(X, ldx, Y, ldy, M, C, C, gamma, beta, eps, stream);
}

extern "C" int edtr_layernorm_padded_bf16(const void* X, int ldx, void* Y, int ldy, int M, int C, int C_real,
                                          const float* gamma, const float* beta, float eps, void* stream) {
  EDTR_REQUIRE(X && Y && gamma && beta, "X/Y/gamma/beta is NULL");
  EDTR_REQUIRE(M > 0 && C > 0 && C % 8 == 0, "bad LayerNorm shape (C %% 8 == 0 required)");
  EDTR_REQUIRE(C_real > 0 && C_real <= C, "C_real (%d) must be in [1, C = %d]", C_real, C);
  EDTR_REQUIRE(C <= 2048, "LayerNorm supports C <= 2048 (got %d)", C);
  EDTR_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C, "bad strides");
  EDTR_REQUIRE(((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y) |
                 reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
               "pointers must be 16-byte aligned");
  const int warps = 8;
  dim3 grid((M + warps - 1) / warps);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(X);
  __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(Y);
  const int vpr = C / 8;
  if (vpr <= 64) EDTR_LAUNCH((layernorm_kernel<2>), grid, warps * 32, 0, st, x, ldx, y, ldy, M, C, C_real, gamma, beta, eps);
  else if (vpr <= 160) EDTR_LAUNCH((layernorm_kernel<5>), grid, warps * 32, 0, st, x, ldx, y, ldy, M, C, C_real, gamma, beta, eps);
  else EDTR_LAUNCH((layernorm_kernel<8>), grid, warps * 32, 0, st, x, ldx, y, ldy, M, C, C_real, gamma, beta, eps);
  return check_launch("layernorm_kernel");
}

extern "C" int edtr_softmax_rows(const float* S, int lds, void* P, int ldp, int M, int N, float scale,
                                 void* stream) {
  EDTR_REQUIRE(S && P && M > 0 && N > 0 && lds >= N && ldp >= N, "bad softmax arguments");
  EDTR_LAUNCH(softmax_rows_kernel, M, 256, 0, static_cast<cudaStream_t>(stream), 
      S, lds, reinterpret_cast<__nv_bfloat16*>(P), ldp, N, scale * 1.4426950408889634f);
  return check_launch("softmax_rows_kernel");
}
