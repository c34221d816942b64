// Flash-style attention for head dim 64 on tcgen05 / TMEM / TMA (sm_100a).
//
// One CTA owns 128 query rows of one (image, head) and streams the keys in tiles of 64 through a 4-stage TMA ring:
// S = Q K^T (UMMA, fp32 in TMEM) -> softmax numerators with one thread per query row (a TMEM lane is a row: no
// shuffles) -> P (packed bf16) back into TMEM -> O += P V (UMMA, A operand from TMEM, V consumed MN-major straight
// from the [keys, 64] TMA box).  O never leaves TMEM inside the loop: the running maximum used for the exponent is
// only raised when the tile maximum exceeds it by more than 2^8 (lazy rescaling), and only then are the O columns of
// the affected rows rescaled in place (tcgen05.ld / st).  Q/K/V are read in place from the projection GEMM outputs
// ([B, L, heads*64] rows): the head split/merge permutes of the reference (model/attention.py:176-203) are TMA
// coordinates.  Kernel history and the measurements behind this structure: profiles/r01g_attention_variants.txt.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace edtr {

constexpr int kAttThreads = 320;   // TMA warp, MMA warp, eight softmax warps
constexpr int kQT = 128;   // query rows per CTA
constexpr int kKT = 64;    // keys per tile
constexpr int kD = 64;     // head dim
constexpr int kTileBytes = kKT * kD * 2;   // 8 KB
constexpr int kAttTmemCols = 256;  // S_A|P_A [0,64) S_B|P_B [64,128) O_A [128,192) O_B [192,256)
constexpr float kRescaleThreshold = 8.f;   // log2 units

// ex2.approx: one MUFU op (exp2f() without -use_fast_math adds range handling around it)
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Packed fp32 pairs (sm_100 FFMA2 / FADD2): one FMA-pipe instruction per two keys.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

struct AttParams {
  void* O;
  int ldo;
  int Lq, Lk;
  float scale_log2;
};

// ------------------------------------------------------------------------------------------------------------------
// Tile-split softmax: the two softmax warp sets do not share a tile.  Set A (warps 2-5) owns the even key
// tiles, set B (warps 6-9) the odd ones; a thread owns one query row of its set's tile (all 64 keys), its own running
// maximum / row sum and its own O accumulator in TMEM (O_A, O_B), so the loop has no cross-warp exchange at all: no
// shared-memory maximum, no named barrier, no P-buffer hand-back.  P (bf16 pairs) overwrites the first 32 columns of
// the S buffer it was computed from and the PV MMA reads it from TMEM; S(j+2) is issued right after PV(j) on the same
// in-order tensor pipe, so the buffer is never overwritten early, and its completion (s_full) also tells the softmax
// set that PV(j) has retired, i.e. that its O accumulator may be rescaled in place.  The two partial results are merged
// once at the end:  O = (O_A 2^(m_A-m) + O_B 2^(m_B-m)) / (l_A 2^(m_A-m) + l_B 2^(m_B-m)).
// kPolyPairs of every 8 key pairs take their exponential from a degree-3 polynomial on the FMA pipe (Cody-Waite range
// reduction with packed f32x2 arithmetic) instead of MUFU.EX2, which is the scarce unit once the loop is this short.
constexpr int kTsStages = 4;
constexpr int kTsSmem = kQT * kD * 2 + kTsStages * 2 * kTileBytes + 1024 + 256;

__device__ __forceinline__ uint64_t add_rm_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t sub_f32x2(uint64_t a, uint64_t b) {   // a - b
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 2^x for a pair, x <= ~2^7: n = floor(x) by the round-down magic add, 2^frac by a cubic (max rel. error 9e-5, below
// bf16 resolution), exponent inserted with one integer multiply-add per element.
__device__ __forceinline__ void exp2_poly_pair(float x0, float x1, float& y0, float& y1) {
  const uint64_t magic = pack_f32x2(12582912.f, 12582912.f);
  const uint64_t x = pack_f32x2(fmaxf(x0, -127.f), fmaxf(x1, -127.f));
  const uint64_t xr = add_rm_f32x2(x, magic);              // low mantissa bits = floor(x)
  const uint64_t fr = sub_f32x2(x, sub_f32x2(xr, magic));  // x - floor(x) in [0, 1)
  uint64_t pl = fma_f32x2(pack_f32x2(0.077119089663028717f, 0.077119089663028717f), fr,
                          pack_f32x2(0.227564394474029541f, 0.227564394474029541f));
  pl = fma_f32x2(pl, fr, pack_f32x2(0.695146143436431885f, 0.695146143436431885f));
  pl = fma_f32x2(pl, fr, pack_f32x2(1.f, 1.f));
  float r0, r1, p0, p1;
  unpack_f32x2(xr, r0, r1);
  unpack_f32x2(pl, p0, p1);
  y0 = __uint_as_float(__float_as_uint(r0) * 8388608u + __float_as_uint(p0));
  y1 = __uint_as_float(__float_as_uint(r1) * 8388608u + __float_as_uint(p1));
}

template <int kPolyPairs>
__global__ void __launch_bounds__(kAttThreads, 2)
attention_ts_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // 16 KB
  uint8_t* sK = sQ + kQT * kD * 2;                 // stages x 8 KB
  uint8_t* sV = sK + kTsStages * kTileBytes;       // stages x 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kTsStages * kTileBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;                    // [stages]
  uint64_t* kv_empty = kv_full + kTsStages;        // [stages]
  uint64_t* s_full = kv_empty + kTsStages;         // [2] S of an even / odd tile is in TMEM
  uint64_t* p_full = s_full + 2;                   // [2] 128 arrivals: P of an even / odd tile is in TMEM
  uint64_t* pv_done = p_full + 2;                  // [2] PV of an even / odd tile has retired
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(pv_done + 2);
  __shared__ float m_ex[2][kQT];
  __shared__ float l_ex[2][kQT];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * kQT;
  const int head = blockIdx.y;
  const int img = blockIdx.z;
  const int nkv = (p.Lk + kKT - 1) / kKT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kTsStages; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&pv_done[i], 1);
      mbar_init(&p_full[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_holder, kAttTmemCols);   // S_A|P_A [0,64) S_B|P_B [64,128) O_A [128,192) O_B [192,256)
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();

  if (warp == 0) {
    const uint32_t sK_u32 = smem_u32(sK), sV_u32 = smem_u32(sV);
    const uint32_t full_u32 = smem_u32(kv_full), empty_u32 = smem_u32(kv_empty);
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, kQT * kD * 2);
      tma_load_4d(sQ, &tmQ, q_full, 0, q0, head, img);
    }
    __syncwarp();
    uint32_t st = 0, ph = 0;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait_u32(empty_u32 + st * 8, ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx_u32(full_u32 + st * 8, 2 * kTileBytes);
        tma_load_4d_u32(sK_u32 + st * kTileBytes, &tmK, full_u32 + st * 8, 0, j * kKT, head, img);
        tma_load_4d_u32(sV_u32 + st * kTileBytes, &tmV, full_u32 + st * 8, 0, j * kKT, head, img);
      }
      __syncwarp();
      if (++st == kTsStages) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc_bf16(kQT, kKT, false);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(kQT, kD, true);
    const uint32_t sQ_u32 = smem_u32(sQ), sK_u32 = smem_u32(sK), sV_u32 = smem_u32(sV);
    const uint32_t full_u32 = smem_u32(kv_full), empty_u32 = smem_u32(kv_empty);
    const uint32_t sfull_u32 = smem_u32(s_full), pfull_u32 = smem_u32(p_full), pvdone_u32 = smem_u32(pv_done);
    uint32_t st_s = 0, ph_s = 0;   // K/V ring position of the next S = Q K^T
    auto issue_s = [&](int j) {
      mbar_wait_u32(full_u32 + st_s * 8, ph_s);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t qdesc = umma_smem_desc_sw128(sQ_u32);
        const uint64_t kdesc = umma_smem_desc_sw128(sK_u32 + st_s * kTileBytes);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_ss(tmem_base + (j & 1) * kKT, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0);
        umma_commit_u32(sfull_u32 + (j & 1) * 8);
      }
      __syncwarp();
      if (++st_s == kTsStages) { st_s = 0; ph_s ^= 1; }
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    if (nkv > 1) issue_s(1);
    uint32_t st = 0;
    for (int j = 0; j < nkv; ++j) {
      const int b = j & 1;
      mbar_wait_u32(pfull_u32 + b * 8, (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t vdesc = umma_smem_desc_sw128(sV_u32 + st * kTileBytes);
#pragma unroll
        for (int k = 0; k < kKT / 16; ++k)   // A = P from TMEM: 16 keys = 8 packed columns; B: +16 rows * 128 B (MN-major V)
          umma_ts(tmem_base + 128 + b * kD, tmem_base + b * kKT + 8 * k, vdesc + 128 * k, idesc_pv, (j >= 2) || (k != 0));
        umma_commit_u32(pvdone_u32 + b * 8);
        umma_commit_u32(empty_u32 + st * 8);
      }
      __syncwarp();
      if (++st == kTsStages) st = 0;
      if (j + 2 < nkv) issue_s(j + 2);   // overwrites S|P buffer b: ordered behind PV(j) on the tensor pipe
    }
  } else {
    const int set = (warp - 2) >> 2;       // 0: even tiles, 1: odd tiles
    const int lg = warp & 3;               // TMEM lane quadrant this warp may access
    const int r = lg * 32 + lane;          // query row inside the tile == TMEM lane
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t s_col = trow + set * kKT;
    const uint32_t o_col = trow + 128 + set * kD;
    float m_used = -INFINITY, l = 0.f;
    const uint64_t sc2 = pack_f32x2(p.scale_log2, p.scale_log2);
    int it = 0;
    for (int j = set; j < nkv; j += 2, ++it) {
      mbar_wait(&s_full[set], it & 1);
      tc_fence_after();
      uint32_t sr[64];
      {
        uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[0]);
        uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&sr[32]);
        tmem_ld32(s_col, lo);
        tmem_ld32(s_col + 32, hi);
      }
      tmem_ld_wait();
      const int valid = p.Lk - j * kKT;
      if (valid < kKT) {  // ragged last tile: mask in place
#pragma unroll
        for (int i = 0; i < kKT; ++i)
          if (i >= valid) sr[i] = 0xff800000u;
      }
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < kKT; i += 4) {
        m0 = fmaxf(m0, __uint_as_float(sr[i]));
        m1 = fmaxf(m1, __uint_as_float(sr[i + 1]));
        m2 = fmaxf(m2, __uint_as_float(sr[i + 2]));
        m3 = fmaxf(m3, __uint_as_float(sr[i + 3]));
      }
      const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      if (it == 0) {
        m_used = mx;
      } else {
        const bool need = (mx - m_used) * p.scale_log2 > kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // O of this set is quiescent: s_full of tile j was committed after PV(j-2)
          const float f = need ? exp2f((m_used - mx) * p.scale_log2) : 1.f;
          if (need) m_used = mx;
          l *= f;
#pragma unroll
          for (int c = 0; c < kD / 16; ++c) {
            uint32_t o[16];
            tmem_ld16(o_col + 16 * c, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st16(o_col + 16 * c, o);
          }
          tmem_st_wait();
        }
      }
      const float mb = m_used * p.scale_log2;
      const uint64_t nmb2 = pack_f32x2(-mb, -mb);
      uint64_t acc0 = 0, acc1 = 0;   // packed (even, odd) partial row sums
#pragma unroll
      for (int c = 0; c < kKT / 16; ++c) {   // 16 keys -> 8 packed bf16 pairs -> 8 TMEM columns of P
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a0, a1;
          unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(sr[16 * c + 2 * i]), __uint_as_float(sr[16 * c + 2 * i + 1])),
                                 sc2, nmb2), a0, a1);
          if (i < kPolyPairs) {
            exp2_poly_pair(a0, a1, a0, a1);
          } else {
            a0 = fast_exp2(a0);
            a1 = fast_exp2(a1);
          }
          if (i & 1) acc1 = add_f32x2(acc1, pack_f32x2(a0, a1));
          else acc0 = add_f32x2(acc0, pack_f32x2(a0, a1));
          pk[i] = pack_bf16(a0, a1);
        }
        tmem_st8(s_col + 8 * c, pk);
      }
      float s0, s1;
      unpack_f32x2(add_f32x2(acc0, acc1), s0, s1);
      l += s0 + s1;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[set]);
    }
    // merge the two partial softmax states of the row (the partner thread lives in warp +-4: same lane quadrant)
    const int n_mine = (nkv - set + 1) >> 1, n_other = (nkv - (set ^ 1) + 1) >> 1;
    m_ex[set][r] = m_used;   // -inf when this set had no tile
    l_ex[set][r] = l;
    named_bar_sync(1 + lg, 64);
    const float m_o = m_ex[set ^ 1][r], l_o = l_ex[set ^ 1][r];
    const float m = fmaxf(m_used, m_o);
    const float f_mine = n_mine > 0 ? exp2f((m_used - m) * p.scale_log2) : 0.f;
    const float f_other = n_other > 0 ? exp2f((m_o - m) * p.scale_log2) : 0.f;
    const float inv = 1.f / (l * f_mine + l_o * f_other);
    const float fA = (set == 0 ? f_mine : f_other) * inv, fB = (set == 0 ? f_other : f_mine) * inv;
    const int nA = (nkv + 1) >> 1, nB = nkv >> 1;
    mbar_wait(&pv_done[0], (nA - 1) & 1);
    if (nB > 0) mbar_wait(&pv_done[1], (nB - 1) & 1);
    tc_fence_after();
    // this thread writes output columns [set*32, set*32+32) of its row
    uint32_t oa[32];
    tmem_ld32(trow + 128 + set * 32, oa);
    tmem_ld_wait();
    float o[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(oa[i]) * fA;
    if (nB > 0) {
      tmem_ld32(trow + 128 + kD + set * 32, oa);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] = fmaf(__uint_as_float(oa[i]), fB, o[i]);
    }
    const int q = q0 + r;
    if (q < p.Lq) {
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.O) +
                          (static_cast<size_t>(img) * p.Lq + q) * p.ldo + head * kD + set * 32;
      uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 u;
        u.x = pack_bf16(o[8 * c + 0], o[8 * c + 1]);
        u.y = pack_bf16(o[8 * c + 2], o[8 * c + 3]);
        u.z = pack_bf16(o[8 * c + 4], o[8 * c + 5]);
        u.w = pack_bf16(o[8 * c + 6], o[8 * c + 7]);
        o4[c] = u;
      }
    }
    tc_fence_before();
  }
  pdl_launch_dependents();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAttTmemCols);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Cross-attention against a short key sequence (the 77 OpenCLIP tokens, model/attention.py:176-203 with context=c_txt).
// The whole K / V of a (sample, head) is 2 x 77 x 64 bf16 = 20 KB, the arithmetic (3.2 GFLOP at 8 x 64 x 64 tokens) is
// far below the size where the TMA / tcgen05 / TMEM pipeline of attention_ts_kernel pays (that kernel spends ~30 us,
// one CTA set-up per 128 query rows for two key tiles), and the launch is bound by streaming Q in and O out.  So: one
// CTA keeps K (row-major) and V (row-major, read through ldmatrix.trans) of its (sample, head) in shared memory and
// walks `chunks` consecutive 128-row query chunks; each of the 8 warps owns 16 query rows per chunk, takes its Q
// fragments straight from global memory (32 contiguous bytes per lane: the k index of Q K^T is permuted identically
// on both operands so that a lane's 16 columns are its own k slots), S = Q K^T and O = P V run on mma.sync m16n8k16,
// the soft-max lives in registers (a row = one quad), no online rescaling (all keys at once).
constexpr int kXK = 80;            // keys held (>= Lk, multiple of 16)
constexpr int kXKPitch = 68;       // K rows: 136 B pitch -> the 8-byte fragment loads of a half-warp hit 16 distinct bank pairs
constexpr int kXVPitch = 72;       // V rows: 144 B pitch -> conflict-free ldmatrix (rows 16-byte aligned)
constexpr int kXWarps = 8;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

__global__ void __launch_bounds__(kXWarps * 32, 2)
cross_attention_small_kernel(const __nv_bfloat16* __restrict__ Q, int ldq, const __nv_bfloat16* __restrict__ K, int ldk,
                             const __nv_bfloat16* __restrict__ V, int ldv, __nv_bfloat16* __restrict__ O, int ldo,
                             int Lq, int Lk, int chunks, float scale_log2) {
  __shared__ __align__(16) __nv_bfloat16 sK[kXK * kXKPitch];
  __shared__ __align__(16) __nv_bfloat16 sV[kXK * kXVPitch];
  const int head = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  pdl_wait();
  const int row_base = blockIdx.x * chunks * (kXWarps * 16);
  // Q fragments of one 16-row slab: lane t owns columns [16 t, 16 t + 16) of rows g and g + 8 (zeros past Lq)
  auto load_q = [&](int q0, uint4 (&q)[4]) {
    q[0] = q[1] = q[2] = q[3] = make_uint4(0u, 0u, 0u, 0u);
    if (q0 + g < Lq) {
      const uint4* src = reinterpret_cast<const uint4*>(Q + (static_cast<size_t>(b) * Lq + q0 + g) * ldq + head * kD + 16 * t);
      q[0] = __ldg(src);
      q[1] = __ldg(src + 1);
    }
    if (q0 + g + 8 < Lq) {
      const uint4* src = reinterpret_cast<const uint4*>(Q + (static_cast<size_t>(b) * Lq + q0 + g + 8) * ldq + head * kD + 16 * t);
      q[2] = __ldg(src);
      q[3] = __ldg(src + 1);
    }
  };
  uint4 qcur[4];
  load_q(row_base + warp * 16, qcur);        // in flight while K / V are staged
  {
    const __nv_bfloat16* kb = K + static_cast<size_t>(b) * Lk * ldk + head * kD;
    const __nv_bfloat16* vb = V + static_cast<size_t>(b) * Lk * ldv + head * kD;
    const uint32_t sK_u32 = smem_u32(sK), sVd_u32 = smem_u32(sV);
    for (int i = tid; i < kXK * 8; i += kXWarps * 32) {     // 80 rows x 8 vectors of 8 bf16, all copies in flight at once
      const int r = i >> 3, c = (i & 7) * 8;
      const int rs = r < Lk ? r : 0;
      const uint32_t n16 = r < Lk ? 16u : 0u, n8 = r < Lk ? 8u : 0u;   // keys >= Lk: zero-filled rows (their probabilities are zero too)
      const __nv_bfloat16* ks = kb + static_cast<size_t>(rs) * ldk + c;
      const __nv_bfloat16* vs = vb + static_cast<size_t>(rs) * ldv + c;
      const uint32_t dk = sK_u32 + (r * kXKPitch + c) * 2;             // 136 B pitch: 8-byte aligned only
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dk), "l"(ks), "r"(n8) : "memory");
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dk + 8), "l"(ks + 4), "r"(n8) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sVd_u32 + (r * kXVPitch + c) * 2), "l"(vs), "r"(n16) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const uint32_t sV_u32 = smem_u32(sV);
  // ldmatrix.x4.trans source row of this lane: matrix m = lane >> 3 -> keys 8 (m & 1) + (lane & 7), d block (m >> 1)
  const uint32_t v_lane = sV_u32 + (((lane >> 3) & 1) * 8 + (lane & 7)) * (kXVPitch * 2) + (lane >> 4) * 16;
  for (int ch = 0; ch < chunks; ++ch) {
    const int q0 = row_base + ch * (kXWarps * 16) + warp * 16;
    if (q0 >= Lq) break;
    const int r0 = q0 + g, r1 = q0 + g + 8;
    const uint32_t qa[8] = {qcur[0].x, qcur[0].y, qcur[0].z, qcur[0].w, qcur[1].x, qcur[1].y, qcur[1].z, qcur[1].w};
    const uint32_t qb[8] = {qcur[2].x, qcur[2].y, qcur[2].z, qcur[2].w, qcur[3].x, qcur[3].y, qcur[3].z, qcur[3].w};
    if (ch + 1 < chunks) load_q(q0 + kXWarps * 16, qcur);   // next slab's Q streams in under this slab's arithmetic
    float s[kXK / 8][4];
#pragma unroll
    for (int j = 0; j < kXK / 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      // k slots of step kk: this lane's columns 16 t + 4 kk + {0, 1} (fragment columns 2t, 2t+1) and + {2, 3} (8 + 2t, ...)
      const uint32_t a[4] = {qa[2 * kk], qb[2 * kk], qa[2 * kk + 1], qb[2 * kk + 1]};
#pragma unroll
      for (int j = 0; j < kXK / 8; ++j) {
        const uint2 kf = *reinterpret_cast<const uint2*>(&sK[(j * 8 + g) * kXKPitch + 16 * t + 4 * kk]);
        mma_bf16_16816(s[j], a, kf.x, kf.y);
      }
    }
    // soft-max over the Lk valid keys; columns 8 j + 2 t + {0, 1} of rows g (s[j][0..1]) and g + 8 (s[j][2..3])
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < kXK / 8; ++j) {
      const int col = j * 8 + 2 * t;
      if (col >= Lk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (col + 1 >= Lk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float m0 = mx0 * scale_log2, m1 = mx1 * scale_log2;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < kXK / 8; ++j) {
      s[j][0] = ex2_approx(fmaf(s[j][0], scale_log2, -m0));
      s[j][1] = ex2_approx(fmaf(s[j][1], scale_log2, -m0));
      s[j][2] = ex2_approx(fmaf(s[j][2], scale_log2, -m1));
      s[j][3] = ex2_approx(fmaf(s[j][3], scale_log2, -m1));
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // O[16 x 64] = P[16 x 80] V[80 x 64]: 5 key steps of 16, 8 d tiles of 8 (two per ldmatrix.x4.trans)
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < kXK / 16; ++kk) {
      const uint32_t a[4] = {pack_bf16(s[2 * kk][0], s[2 * kk][1]), pack_bf16(s[2 * kk][2], s[2 * kk][3]),
                             pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]), pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t vf[4];
        ldmatrix_x4_trans(vf, v_lane + kk * 16 * (kXVPitch * 2) + np * 32);
        mma_bf16_16816(o[2 * np], a, vf[0], vf[1]);
        mma_bf16_16816(o[2 * np + 1], a, vf[2], vf[3]);
      }
    }
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    if (r0 < Lq) {
      __nv_bfloat16* dst = O + (static_cast<size_t>(b) * Lq + r0) * ldo + head * kD + 2 * t;
#pragma unroll
      for (int n = 0; n < 8; ++n) *reinterpret_cast<uint32_t*>(dst + n * 8) = pack_bf16(o[n][0] * i0, o[n][1] * i0);
    }
    if (r1 < Lq) {
      __nv_bfloat16* dst = O + (static_cast<size_t>(b) * Lq + r1) * ldo + head * kD + 2 * t;
#pragma unroll
      for (int n = 0; n < 8; ++n) *reinterpret_cast<uint32_t*>(dst + n * 8) = pack_bf16(o[n][2] * i1, o[n][3] * i1);
    }
  }
  pdl_launch_dependents();
}

// EDTR_XATTN_SMALL=0 sends the short-key launches to attention_ts_kernel again (A/B switch).
static bool cross_attention_small_enabled() {
  static const bool v = [] {
    const char* e = getenv("EDTR_XATTN_SMALL");
    return e == nullptr || e[0] != '0';
  }();
  return v;
}

int prime_attention_attributes() {
  cudaError_t e = cudaFuncSetAttribute(attention_ts_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmem);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(attention_ts_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTsSmem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(attention): %s", cudaGetErrorString(e));
    return EDTR_ERR_CUDA;
  }
  return EDTR_OK;
}

// EDTR_ATT_POLY=0 computes every exponential on the MUFU unit; the default moves 2 of every 8 key pairs to the
// FMA-pipe polynomial (measured 2.5 % faster on the L = 4096 self-attention).
static bool attention_poly() {
  static const bool v = [] {
    const char* e = getenv("EDTR_ATT_POLY");
    return e == nullptr || e[0] != '0';
  }();
  return v;
}

static int make_head_tmap(CUtensorMap* tm, const void* base, int ld, int B, int heads, int L, int box_rows) {
  uint64_t dims[4] = {static_cast<uint64_t>(kD), static_cast<uint64_t>(L), static_cast<uint64_t>(heads),
                      static_cast<uint64_t>(B)};
  uint64_t strides[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(kD) * 2,
                         static_cast<uint64_t>(ld) * 2 * L};
  uint32_t box[4] = {kD, static_cast<uint32_t>(box_rows), 1, 1};
  return make_tmap_bf16(tm, base, 4, dims, strides, box);
}

}  // namespace edtr

using namespace edtr;

extern "C" int edtr_attention_bf16(const void* Q, int ldq, const void* K, int ldk, const void* V, int ldv,
                                   void* O, int ldo, int B, int heads, int Lq, int Lk, float scale,
                                   void* stream) {
  EDTR_REQUIRE(Q && K && V && O, "Q/K/V/O is NULL");
  EDTR_REQUIRE(B > 0 && heads > 0 && Lq > 0 && Lk > 0, "bad attention shape");
  EDTR_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "row strides must be multiples of 8");
  EDTR_REQUIRE(ldq >= heads * kD && ldk >= heads * kD && ldv >= heads * kD && ldo >= heads * kD,
               "row strides must cover heads*64 columns");
  EDTR_REQUIRE(((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(K) | reinterpret_cast<uintptr_t>(V) |
                 reinterpret_cast<uintptr_t>(O)) & 15) == 0, "Q/K/V/O must be 16-byte aligned");
  EDTR_REQUIRE(heads <= 65535 && B <= 65535, "grid too large");
  if (Lk <= kXK && cross_attention_small_enabled()) {
    // chunks of 128 query rows per CTA: at most two — measured on B200 inside the sample graph (profiles/r03h): 1 / 2 /
    // one-wave (5, 3, 2 on the three levels) / 8 chunks per CTA give 47.8 / 47.2 / 47.6 / 48.0 ms per 4-step sample; a
    // few waves of short CTAs hide the Q latency better than one wave of long ones, K / V reloads are L2 hits
    const int nchunk = (Lq + kXWarps * 16 - 1) / (kXWarps * 16);
    int chunks = static_cast<int>((static_cast<long long>(nchunk) * heads * B + 2 * 148 - 1) / (2 * 148));
    if (chunks > 2) chunks = 2;
    static const int chunks_override = [] {      // EDTR_XATTN_CHUNKS=<n>: A/B switch for the chunk count per CTA
      const char* e = getenv("EDTR_XATTN_CHUNKS");
      return e == nullptr ? 0 : atoi(e);
    }();
    if (chunks_override > 0) chunks = chunks_override;
    if (chunks < 1) chunks = 1;
    if (chunks > 8) chunks = 8;
    dim3 grid((nchunk + chunks - 1) / chunks, heads, B);
    EDTR_LAUNCH(cross_attention_small_kernel, grid, kXWarps * 32, 0, static_cast<cudaStream_t>(stream),
                reinterpret_cast<const __nv_bfloat16*>(Q), ldq, reinterpret_cast<const __nv_bfloat16*>(K), ldk,
                reinterpret_cast<const __nv_bfloat16*>(V), ldv, reinterpret_cast<__nv_bfloat16*>(O), ldo, Lq, Lk, chunks,
                scale * 1.4426950408889634f);
    return check_launch("cross_attention_small_kernel");
  }
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_head_tmap(&tmQ, Q, ldq, B, heads, Lq, kQT);
  if (rc) return rc;
  rc = make_head_tmap(&tmK, K, ldk, B, heads, Lk, kKT);
  if (rc) return rc;
  rc = make_head_tmap(&tmV, V, ldv, B, heads, Lk, kKT);
  if (rc) return rc;
  AttParams p;
  p.O = O; p.ldo = ldo; p.Lq = Lq; p.Lk = Lk;
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((Lq + kQT - 1) / kQT, heads, B);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (attention_poly())
    EDTR_LAUNCH(attention_ts_kernel<2>, grid, kAttThreads, kTsSmem, st, tmQ, tmK, tmV, p);
  else
    EDTR_LAUNCH(attention_ts_kernel<0>, grid, kAttThreads, kTsSmem, st, tmQ, tmK, tmV, p);
  return check_launch("attention_ts_kernel");
}
