// Fused tensor-core kernel: q(z|u,s) chain + pairwise distances + group reduction, 4 <= S <= 128.
//
// The S x L tile of counterfactual representations of a cell never leaves the SM (north_star):
// after GEMM5 the residuals (z_base cancels in every distance, reference _module.py:166-170, 331) are
// centred per cell, split into (hi, lo) f16 and fed back to the tensor cores as BOTH operands of a
// Gram product  G = Zh Zh^T + Zh Zl^T + Zl Zh^T  (fp32 accumulate in TMEM).  Row r of G gives
// d[r][j] = sqrt(n_r + n_j - 2 G[r][j]) with fp32 norms; pairs whose squared distance is cancellation
// dominated (< 2^-12 of n_r + n_j; the Gram error is ~2^-20 of it) are recomputed by direct
// differences, the diagonal is forced to zero.  Rows are stored per cell (keep_cell) and/or added to
// a [128 x 128] fp32 accumulator that lives in TMEM while consecutive tiles carry the same composite
// group code (cells arrive sorted by it) and is flushed with float64 atomics when the code changes.
//
// Shape of the kernel (what the ncu profile of the first version asked for: its fully unrolled
// 718 KB body spent 48 % of the samples in instruction-fetch stalls with 2 warps per scheduler):
//   * one CTA per SM, 512 threads = 2 warpgroups x 256; each warpgroup owns one 128-row tile in
//     flight (own A buffers, own 160 TMEM columns, own mbarrier); the 50 KB weight image and the TMEM
//     accumulator are shared (the accumulator behind a shared-memory lock, taken once per tile);
//   * TWO threads per row: thread (r, half) owns the column half `half` of every epilogue (the
//     LayerNorm partial sums are exchanged through shared memory), so a tile's latency halves;
//   * every epilogue is a loop over 32-column chunks (re-loaded from TMEM), which keeps the hot code
//     in the instruction cache and the kernel at <= 128 registers.
// The 128-wide hidden activations are written back in place into TMEM, per 32-column chunk as
// [16 words hi | 16 words lo] (f16 pairs), and read by the next GEMM as its A operand (TS-mode MMA).
#pragma once
#include "tc_kernels.cuh"

namespace mrvi {

constexpr int kFusedThreads = 512;

#ifdef MRVI_DBG_FUSED
// clock64 timeline of CTA 0 (tools/dbg/run_dbgf.py): [warpgroup][tile < 32][mark < 32], thread 0 of each warpgroup
__device__ long long g_dbg_fused[2 * 32 * 32];
#define FT(i_) do { if (blockIdx.x == 0 && wt == 0 && dbg_tile < 32) g_dbg_fused[(wg * 32 + dbg_tile) * 32 + (i_)] = clock64(); } while (0)
#else
#define FT(i_) do { } while (0)
#endif

struct FusedSmemMap {
  uint32_t w;            // weight image (hi | lo)
  uint32_t wg[2];        // per-warpgroup scratch: [a1h a1l a3h a3l] overlaid later by [zf] (z output only), then [zh zl | T | cref]
  uint32_t off_a1l, off_a3h, off_a3l;   // offsets inside the scratch (a1h at 0); a1l holds the 4 feature chunks only
  uint32_t off_zf, off_zh, off_zl, off_T, off_cref;  // overlay offsets
  uint32_t nrm[2];       // [128] fp32
  uint32_t xch[2];       // [2 halves][128] float2: partial row statistics
  uint32_t lnv;          // TcLn copy (640 floats)
  uint32_t ctrl;         // barriers, tmem slot, lock, run state
  uint32_t total;
  int zf_ld;             // row stride (floats) of the fp32 z tile, odd
  int Kz;                // padded K of the Gram operands (32 or 64)
  int cref_rows;         // reference rows per cell whose mean is the centre of the Gram operands (2; 1 when n_latent > 32)
};

__host__ __device__ inline FusedSmemMap fused_smem_map(uint32_t wimg_bytes, int K1, int K3, int L) {
  FusedSmemMap m;
  uint32_t o = 0;
  auto up = [](uint32_t b) { return (b + 1023u) & ~1023u; };
  m.w = o; o += up(wimg_bytes);
  const uint32_t a1 = up(128 * K1 * 2), a1l = 128 * 32 * 2, a3 = up(128 * K3 * 2);
  m.off_a1l = a1; m.off_a3h = a1 + a1l; m.off_a3l = a1 + a1l + a3;
  const uint32_t abytes = a1 + a1l + 2 * a3;
  m.Kz = L <= 32 ? 32 : 64;
  m.zf_ld = m.Kz + 1;
  // A1 / A3 are dead after GEMM5.  The fp32 z tile (only when z is an output) is copied out before the Gram operands
  // overwrite it; behind the operands: the keep_cell staging tiles T (8 warps x 2 KB) and the per-cell reference rows
  // cref [32 cells][2][Kz] fp32 (samples 0 and S / 2 of each cell; their mean is subtracted before the Gram).
  const uint32_t zf = up(128 * m.zf_ld * 4), zk = up(128 * m.Kz * 2);
  m.off_zf = 0; m.off_zh = 0; m.off_zl = zk; m.off_T = 2 * zk; m.off_cref = 2 * zk + 16384;
  uint32_t scratch = abytes;
  if (scratch < zf) scratch = zf;
  m.cref_rows = m.Kz == 32 ? 2 : 1;
  if (scratch < m.off_cref + 32u * m.Kz * 4u * m.cref_rows) scratch = m.off_cref + 32u * m.Kz * 4u * m.cref_rows;
  m.wg[0] = o; o += scratch;
  m.wg[1] = o; o += scratch;
  m.nrm[0] = o; o += 512;
  m.nrm[1] = o; o += 512;
  m.xch[0] = o; o += 2048;
  m.xch[1] = o; o += 2048;
  m.lnv = o; o += (uint32_t)((sizeof(TcLn) + 127) & ~size_t(127));
  m.ctrl = o; o += 128;
  m.total = o;
  return m;
}

struct FusedCtrl {
  uint64_t bar_w;
  uint64_t bar_mma[2];
  uint32_t tmem_slot;
  int lock;
  int cur_code;       // composite group code currently held in the TMEM accumulator
  int cur_valid;
  long long cur_cell; // a cell of that run (to look up its per-key group ids at flush time)
};

__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, 256;" ::"r"(wg + 1) : "memory"); }
// the two warps that own the column halves of the same 32 rows (warps w and w + 4 of a warpgroup: same scheduler, same TMEM
// lane quarter): the LayerNorm statistics exchange needs only them, not the whole warpgroup
__device__ __forceinline__ void pair_sync(int wg, int r) { asm volatile("bar.sync %0, 64;" ::"r"(3 + wg * 4 + (r >> 5)) : "memory"); }
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float h2f_lo(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))); }
__device__ __forceinline__ float h2f_hi(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }

// ---- packed fp32x2 helpers (FFMA2 / FMUL2 / FADD2 halve the issue slots of the fp32 epilogue math)
// TWICE gelu(tanh): x + x tanh(...).  The factor 1/2 is folded into the rows of the next GEMM's weight operand that
// multiply this activation (tc_build: exact, a power of two), which saves one packed multiply per pair of values in
// every epilogue of the chain.
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  const float2 c1 = make_float2(0.7978845608028654f, 0.7978845608028654f);
  const float2 c2 = make_float2(0.7978845608028654f * 0.044715f, 0.7978845608028654f * 0.044715f);
  const float2 u = __fmul2_rn(x, x);
  const float2 p = __ffma2_rn(u, c2, c1);
  const float2 y = __fmul2_rn(x, p);
  const float2 t = make_float2(tanh_approx(y.x), tanh_approx(y.y));
  return __ffma2_rn(x, t, x);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// Exact direct-difference distances for the cancellation-dominated entries of a warp's rows, computed by the WHOLE
// warp (a lane-serial version keeps 31 lanes, and everybody behind the next barrier, waiting for thousands of
// cycles): every lane passes its operand row and the bit mask of its flagged entries among columns
// [jbase, jbase + 32); for each flagged (lane, q) the 32 lanes split the K range, warp-reduce the squared
// differences, and the owner lane replaces g[q].  Operands: K-major (hi, lo) f16 buffers, chunk stride LBO bytes.
// Must be called by all 32 lanes.
template <uint32_t LBO>
__device__ __noinline__ float exact_dist_warp(const uint8_t* zh, const uint8_t* zl, int a, int b, int L) {
  const int lane = threadIdx.x & 31;
  float e = 0.f;
  for (int k = lane; k < L; k += 32) {
    const uint32_t ko = (uint32_t)(k >> 3) * LBO + (uint32_t)(k & 7) * 2u;
    const float za = __half2float(*reinterpret_cast<const __half*>(zh + ko + a * 16)) +
                     __half2float(*reinterpret_cast<const __half*>(zl + ko + a * 16));
    const float zb = __half2float(*reinterpret_cast<const __half*>(zh + ko + b * 16)) +
                     __half2float(*reinterpret_cast<const __half*>(zl + ko + b * 16));
    const float df = zb - za;
    e = fmaf(df, df, e);
  }
  return sqrtf(warp_sum(e));
}
// (inlined: g stays in registers; only the reduction above is an out-of-line call)
template <uint32_t LBO>
__device__ __forceinline__ void exact_fix_warp(float (&g)[32], uint32_t flagged, const uint8_t* zh, const uint8_t* zl, int row,
                                               int jbase, int L) {
  const int lane = threadIdx.x & 31;
  uint32_t any = __ballot_sync(0xffffffffu, flagged != 0u);
#pragma unroll 1
  while (any) {
    const int leader = __ffs(any) - 1;
    any &= any - 1;
    uint32_t fl = __shfl_sync(0xffffffffu, flagged, leader);
    const int a = __shfl_sync(0xffffffffu, row, leader);
#pragma unroll 1
    while (fl) {
      const int q = __ffs(fl) - 1;
      fl &= fl - 1;
      const float d = exact_dist_warp<LBO>(zh, zl, a, jbase + q, L);
      const bool mine = lane == leader;
#pragma unroll
      for (int qq = 0; qq < 32; ++qq)
        if (mine && qq == q) g[qq] = d;
    }
  }
}

// 32 Gram entries of one row -> distances, in place; returns min over the entries of (d^2 - 2^-12 (n_i + n_j)):
// negative = some entry is cancellation dominated (the caller recomputes those exactly).  DIAG: entry q == lane is
// the row's diagonal entry (forced to zero, excluded from the test).
template <bool DIAG, int THR_LOG2 = 12>
__device__ __forceinline__ float dist_chunk_fast(float (&g)[32], const float* nrm_j, float n_i, int lane) {
  constexpr float kThr = -1.0f / (float)(1 << THR_LOG2);
  const float2 ni2 = make_float2(n_i, n_i), m2 = make_float2(-2.0f, -2.0f), thr2 = make_float2(kThr, kThr);
  float tmin = 0.f;
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4) {
    const float4 nj = *reinterpret_cast<const float4*>(nrm_j + 4 * q4);
    const float2 njp[2] = {make_float2(nj.x, nj.y), make_float2(nj.z, nj.w)};
#pragma unroll
    for (int e2 = 0; e2 < 2; ++e2) {
      const int q = 4 * q4 + 2 * e2;
      const float2 sn = __fadd2_rn(ni2, njp[e2]);
      const float2 d2 = __ffma2_rn(make_float2(g[q], g[q + 1]), m2, sn);
      const float2 tt = __ffma2_rn(sn, thr2, d2);  // < 0: cancellation dominated
      float t0 = tt.x, t1 = tt.y;
      // |d2|: the operand modifier is free; an entry with d2 < 0 is below the threshold (tt < 0) and recomputed anyway
      float r0 = sqrt_approx(fabsf(d2.x)), r1 = sqrt_approx(fabsf(d2.y));
      if (DIAG) {
        const bool dg0 = q == lane, dg1 = q + 1 == lane;
        t0 = dg0 ? 0.f : t0; t1 = dg1 ? 0.f : t1;
        r0 = dg0 ? 0.f : r0; r1 = dg1 ? 0.f : r1;
      }
      tmin = fminf(tmin, fminf(t0, t1));
      g[q] = r0; g[q + 1] = r1;
    }
  }
  return tmin;
}

// LayerNorm + gelu of the 128-column accumulator row, this thread's 64-column half, written back in
// place per 32-column chunk as [16 hi words | 16 lo words].  g_s / b_s: shared-memory LN vectors.
__device__ __forceinline__ void ep128(uint32_t t_lane, int half, int r, int wg, const float* g_s, const float* b_s,
                                      float2* xch) {
  float2 a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f);
  {
    float v[32], w[32];   // both chunks travel together: one TMEM round trip for the statistics pass
    tmem_ld32(t_lane + (2 * half) * 32, v);
    tmem_ld32(t_lane + (2 * half + 1) * 32, w);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 x = make_float2(v[2 * i], v[2 * i + 1]), y = make_float2(w[2 * i], w[2 * i + 1]);
      a1 = __fadd2_rn(a1, x);
      a2 = __ffma2_rn(x, x, a2);
      a1 = __fadd2_rn(a1, y);
      a2 = __ffma2_rn(y, y, a2);
    }
  }
  float s1 = a1.x + a1.y, s2 = a2.x + a2.y;
  xch[half * 128 + r] = make_float2(s1, s2);
  pair_sync(wg, r);
  const float2 o = xch[(half ^ 1) * 128 + r];
  s1 += o.x; s2 += o.y;
  const float mean = s1 * (1.0f / 128);
  const float var = fmaxf(fmaf(s2, 1.0f / 128, -mean * mean), 0.f);
  const float rstd = rsqrt_approx(var + kLnEps);
  const float2 rs2 = make_float2(rstd, rstd), nb2 = make_float2(-mean * rstd, -mean * rstd);
#pragma unroll 1
  for (int pp = 0; pp < 2; ++pp) {
    const int p = 2 * half + pp;
    float v[32];
    tmem_ld32(t_lane + p * 32, v);
    tmem_ld_wait();
    uint32_t out[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 g4 = *reinterpret_cast<const float4*>(g_s + p * 32 + 4 * q);
      const float4 b4 = *reinterpret_cast<const float4*>(b_s + p * 32 + 4 * q);
      float2 x0 = __ffma2_rn(make_float2(v[4 * q], v[4 * q + 1]), rs2, nb2);
      float2 x1 = __ffma2_rn(make_float2(v[4 * q + 2], v[4 * q + 3]), rs2, nb2);
      x0 = gelu_fast2(__ffma2_rn(x0, make_float2(g4.x, g4.y), make_float2(b4.x, b4.y)));
      x1 = gelu_fast2(__ffma2_rn(x1, make_float2(g4.z, g4.w), make_float2(b4.z, b4.w)));
      split2v(x0, out[2 * q], out[16 + 2 * q]);
      split2v(x1, out[2 * q + 1], out[16 + 2 * q + 1]);
    }
    tmem_st32(t_lane + p * 32, out);
  }
  tmem_st_wait();
}

// LayerNorm + gelu of the 32-column accumulator row, this thread's 16 columns -> two 16-byte K chunks
// (hi and lo) of the next A operand in shared memory.
__device__ __forceinline__ void ep32(uint32_t t_lane32, int half, int r, int wg, const float* g_s, const float* b_s,
                                     float2* xch, uint8_t* row_h, uint8_t* row_l) {
  float v[16];
  tmem_ld16(t_lane32 + half * 16, v);
  tmem_ld_wait();
  float2 a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 x = make_float2(v[2 * i], v[2 * i + 1]);
    a1 = __fadd2_rn(a1, x);
    a2 = __ffma2_rn(x, x, a2);
  }
  float s1 = a1.x + a1.y, s2 = a2.x + a2.y;
  xch[half * 128 + r] = make_float2(s1, s2);
  pair_sync(wg, r);
  const float2 o = xch[(half ^ 1) * 128 + r];
  s1 += o.x; s2 += o.y;
  const float mean = s1 * (1.0f / 32);
  const float var = fmaxf(fmaf(s2, 1.0f / 32, -mean * mean), 0.f);
  const float rstd = rsqrt_approx(var + kLnEps);
  const float2 rs2 = make_float2(rstd, rstd), nb2 = make_float2(-mean * rstd, -mean * rstd);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i0 = c * 8 + 2 * j;
      const float2 gg = *reinterpret_cast<const float2*>(g_s + half * 16 + i0);
      const float2 bb = *reinterpret_cast<const float2*>(b_s + half * 16 + i0);
      float2 x = __ffma2_rn(make_float2(v[i0], v[i0 + 1]), rs2, nb2);
      x = gelu_fast2(__ffma2_rn(x, gg, bb));
      split2v(x, h[j], l[j]);
    }
    *reinterpret_cast<uint4*>(row_h + (half * 2 + c) * 2048) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(row_l + (half * 2 + c) * 2048) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// LayerNorm + gelu of the 32-wide accumulator stored as two halves [A.W_hi | A.W_lo] (columns [0,32) / [32,64) of the D2
// region, see TcOp): this thread's 16 columns -> two 16-byte K chunks (hi and lo) of the next A operand in shared memory.
__device__ __forceinline__ void ep32x(uint32_t t_d2_lane, int half, int r, int wg, const float* g_s, const float* b_s, float2* xch,
                                      uint8_t* row_h, uint8_t* row_l) {
  float v[16], w[16];
  tmem_ld16(t_d2_lane + half * 16, v);
  tmem_ld16(t_d2_lane + 32 + half * 16, w);
  tmem_ld_wait();
  float2 a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 x = __fadd2_rn(make_float2(v[2 * i], v[2 * i + 1]), make_float2(w[2 * i], w[2 * i + 1]));
    v[2 * i] = x.x; v[2 * i + 1] = x.y;
    a1 = __fadd2_rn(a1, x);
    a2 = __ffma2_rn(x, x, a2);
  }
  float s1 = a1.x + a1.y, s2 = a2.x + a2.y;
  xch[half * 128 + r] = make_float2(s1, s2);
  pair_sync(wg, r);
  const float2 o = xch[(half ^ 1) * 128 + r];
  s1 += o.x; s2 += o.y;
  const float mean = s1 * (1.0f / 32);
  const float var = fmaxf(fmaf(s2, 1.0f / 32, -mean * mean), 0.f);
  const float rstd = rsqrt_approx(var + kLnEps);
  const float2 rs2 = make_float2(rstd, rstd), nb2 = make_float2(-mean * rstd, -mean * rstd);
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i0 = c * 8 + 2 * j;
      const float2 gg = *reinterpret_cast<const float2*>(g_s + half * 16 + i0);
      const float2 bb = *reinterpret_cast<const float2*>(b_s + half * 16 + i0);
      float2 x = __ffma2_rn(make_float2(v[i0], v[i0 + 1]), rs2, nb2);
      x = gelu_fast2(__ffma2_rn(x, gg, bb));
      split2v(x, h[j], l[j]);
    }
    *reinterpret_cast<uint4*>(row_h + (half * 2 + c) * 2048) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(row_l + (half * 2 + c) * 2048) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// [128 x 128] (A in TMEM, chunked [hi16 | lo16] layout) x concatenated [128 x (32 | 32)]: T_hi.[W_hi | W_lo] as ONE N = 64
// instruction per K slice + T_lo.W_hi (N = 32): 16 MMAs instead of 24 (an MMA costs ~42 clk for N <= 64 whatever N)
__device__ __forceinline__ void issue_gemm_ts2_cat(uint32_t d_tmem, uint32_t a_tmem, uint32_t s_w, const TcOp& b, uint32_t idesc64, uint32_t idesc32) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint64_t bd = umma_desc(s_w + b.h + ks * 2 * b.lbo, b.lbo, 128);
    const uint32_t a = a_tmem + (ks >> 1) * 32 + (ks & 1) * 8;
    umma_f16_ts(d_tmem, a, bd, idesc64, 1u);   // accumulates onto the skip part (issue_gemm_cat, issued earlier, clears)
    umma_f16_ts(d_tmem, a + 16, bd, idesc32, 1u);
  }
}
// [128 x K] (hi, lo in shared memory) x concatenated [K x (32 | 32)]: A_hi.[W_hi | W_lo] + A_lo.W_hi; clears D.  This is the
// skip-connection part of GEMM2 / GEMM4: it does not depend on the epilogue in front of it, so it is issued right behind
// GEMM1 / GEMM3 and runs under epilogue 1 / 3 (6 MMAs off the critical path of each)
__device__ __forceinline__ void issue_gemm_cat(uint32_t d_tmem, uint32_t a_h, uint32_t a_l, uint32_t s_w, const TcOp& b, int K, int K_lo,
                                               uint32_t idesc64, uint32_t idesc32) {
  const uint32_t a_lbo = 128 * 16;
  for (int ks = 0; ks < K / 16; ++ks) {
    const uint64_t bd = umma_desc(s_w + b.h + ks * 2 * b.lbo, b.lbo, 128);
    umma_f16(d_tmem, umma_desc(a_h + ks * 2 * a_lbo, a_lbo, 128), bd, idesc64, ks == 0 ? 0u : 1u);
    if (ks * 16 < K_lo) umma_f16(d_tmem, umma_desc(a_l + ks * 2 * a_lbo, a_lbo, 128), bd, idesc32, 1u);
  }
}

// [128 x 128] (A in TMEM, chunked [hi16|lo16] layout) x [128 x N], weights (hi, lo): Ah.Wh + Al.Wh + Ah.Wl
__device__ __forceinline__ void issue_gemm_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t s_w, const TcOp& b, uint32_t idesc) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint64_t bd = umma_desc(s_w + b.h + ks * 2 * b.lbo, b.lbo, 128);
    const uint64_t bl = umma_desc(s_w + b.l + ks * 2 * b.lbo, b.lbo, 128);
    const uint32_t a = a_tmem + (ks >> 1) * 32 + (ks & 1) * 8;
    umma_f16_ts(d_tmem, a, bd, idesc, ks == 0 ? 0u : 1u);
    umma_f16_ts(d_tmem, a + 16, bd, idesc, 1u);
    umma_f16_ts(d_tmem, a, bl, idesc, 1u);
  }
}

// flush this thread's two chunks of its accumulator lane into the float64 group sums and clear them
__device__ __forceinline__ void flush_acc(uint32_t t_acc_lane, int half, const GroupPlan& gp, long long rep_cell, int S,
                                          int c0, bool lane_valid, int r) {
#pragma unroll 1
  for (int pp = 0; pp < 2; ++pp) {
    const int c = 2 * half + pp;
    float a[32];
    tmem_ld32(t_acc_lane + c * 32, a);
    tmem_ld_wait();
    if (lane_valid) {
#pragma unroll 1
      for (int k = 0; k < gp.n_keys; ++k) {
        const int g = gp.codes[k][rep_cell];
        if (g < 0 || g >= gp.n_groups[k]) continue;
        double* base = gp.sums[k] + ((size_t)g * S + (r - c0)) * S;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int j = c * 32 + q - c0;
          if (j >= 0 && j < S && j != r - c0) atomicAdd(base + j, (double)a[q]);
        }
      }
    }
    uint32_t z[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) z[q] = 0u;
    tmem_st32(t_acc_lane + c * 32, z);
  }
  tmem_st_wait();
}

__global__ void __launch_bounds__(kFusedThreads, 1)
k_chain_fused_tc(const __grid_constant__ TcNet tc, CellBuf cb, int64_t n_cells, int S, int L, int Lq,
                 float* __restrict__ z_out, float* __restrict__ cell_out, GroupPlan gp, int64_t n_tiles,
                 int cells_per_tile, int64_t tiles_per_cta, int tiles_per_cell) {
  // no-swizzle UMMA operands and bulk copies need 16-byte alignment only; keeping the plain array (no integer
  // re-alignment) lets the compiler see the shared address space (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem_tc[];
  uint8_t* const smem = smem_tc;
  const FusedSmemMap sm = fused_smem_map(tc.wimg_bytes, tc.K1, tc.K3, L);
  FusedCtrl* ctrl = reinterpret_cast<FusedCtrl*>(smem + sm.ctrl);
  const int tid = threadIdx.x, wg = tid >> 8, wt = tid & 255, half = wt >> 7, r = wt & 127, warp_q = (wt >> 5) & 3;
  const int Kz = sm.Kz, zld = sm.zf_ld;
  const bool kc2 = sm.cref_rows == 2;
  float* lnv = reinterpret_cast<float*>(smem + sm.lnv);

  if (tid == 0) {
    mbar_init(&ctrl->bar_w, 1);
    mbar_init(&ctrl->bar_mma[0], 1);
    mbar_init(&ctrl->bar_mma[1], 1);
    ctrl->lock = 0; ctrl->cur_code = -1; ctrl->cur_valid = 0; ctrl->cur_cell = 0;
    fence_mbar_init();
  }
  for (int i = tid; i < (int)(sizeof(TcLn) / sizeof(float)); i += kFusedThreads) lnv[i] = reinterpret_cast<const float*>(&tc.ln)[i];
  if (tid < 32) tmem_alloc(&ctrl->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(&ctrl->bar_w, tc.wimg_bytes);
    bulk_g2s(smem + sm.w, tc.wimg, tc.wimg_bytes, &ctrl->bar_w);
  }
  const float *g1 = lnv, *b1 = lnv + 128, *g2 = lnv + 256, *b2 = lnv + 288, *g3 = lnv + 320, *b3 = lnv + 448,
              *g4 = lnv + 576, *b4 = lnv + 608;
  // TMEM columns: warpgroup g: [g*192, +128) = D1/T/D3/T/D5/Gram, [g*192+128, +64) = D2/D4 as [A.W_hi | A.W_lo]; accumulator [384, 512)
  const uint32_t t_r0 = tmem_base + wg * 192, t_d2 = t_r0 + 128;
  const uint32_t lane_off = (uint32_t)(warp_q * 32) << 16;
  const uint32_t t_lane = t_r0 + lane_off, t_d2_lane = t_d2 + lane_off, t_acc_lane = tmem_base + 384 + lane_off;
  uint8_t* scratch = smem + sm.wg[wg];
  const uint32_t s_w = smem_u32(smem + sm.w);
  uint8_t* row_a1h = scratch + r * 16;
  uint8_t* row_a1l = scratch + sm.off_a1l + r * 16;
  uint8_t* row_a3h = scratch + sm.off_a3h + r * 16;
  uint8_t* row_a3l = scratch + sm.off_a3l + r * 16;
  float* zf = reinterpret_cast<float*>(scratch + sm.off_zf);
  uint8_t* zh = scratch + sm.off_zh;
  uint8_t* zl = scratch + sm.off_zl;
  float* cref_s = reinterpret_cast<float*>(scratch + sm.off_cref);
  float* nrm_s = reinterpret_cast<float*>(smem + sm.nrm[wg]);
  float2* xch = reinterpret_cast<float2*>(smem + sm.xch[wg]);
  uint64_t* bar_mma = &ctrl->bar_mma[wg];
  const uint32_t idesc128 = umma_idesc_f16(128), idesc64 = umma_idesc_f16(64), idesc32 = umma_idesc_f16(32), idescN5 = umma_idesc_f16(tc.N5);
  uint32_t phase = 0;
  const bool need_dist = (cell_out != nullptr) || (gp.n_keys > 0);
  const bool issuer = wt == 0;   // lock / run-state bookkeeping of the warpgroup
  // MMA issue: the first warp of each warpgroup, as a warp-uniform branch with elect.sync inside (see elect_one); every
  // address the issue code uses is derived from the shuffled (= provably uniform) warp index
  const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const bool issue_warp = (warp_u & 7) == 0;
  const uint32_t iu_tr0 = __shfl_sync(0xffffffffu, tmem_base, 0) + (uint32_t)(warp_u >> 3) * 192u, iu_td2 = iu_tr0 + 128u;
  const uint32_t iu_a1h = smem_u32(smem + sm.wg[warp_u >> 3]), iu_a1l = iu_a1h + sm.off_a1l, iu_a3h = iu_a1h + sm.off_a3h,
                 iu_a3l = iu_a1h + sm.off_a3l, iu_zh = iu_a1h + sm.off_zh, iu_zl = iu_a1h + sm.off_zl;
  uint64_t* iu_bar = &ctrl->bar_mma[warp_u >> 3];

  // static row -> (local cell, sample) map of this thread
  const int lc = r / S, s = r - lc * S, c0 = lc * S;
  const bool row_in_tile = lc < cells_per_tile;
  // warp-uniform Gram column chunks this warp needs (rows of a warp may straddle cells)
  int cmin = row_in_tile ? c0 : 128, cmax = row_in_tile ? c0 + S : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
    cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
  }
  const int ch_lo = max(cmin >> 5, 2 * half), ch_hi = min((cmax + 31) >> 5, 2 * half + 2);  // this thread's chunks

  if (gp.n_keys > 0 && wg == 0) {  // clear the accumulator once
    uint32_t z[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) z[q] = 0u;
    tmem_st32(t_acc_lane + (2 * half) * 32, z);
    tmem_st32(t_acc_lane + (2 * half + 1) * 32, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  mbar_wait(&ctrl->bar_w, 0);
  const float a_h = half ? tc.alpha[1] : tc.alpha[0], b_h = half ? tc.beta[1] : tc.beta[0];

  const int64_t tile_begin = (int64_t)blockIdx.x * tiles_per_cta;
  const int64_t tile_end = min(n_tiles, tile_begin + tiles_per_cta);
  for (int64_t tile = tile_begin + wg; tile < tile_end; tile += 2) {
#ifdef MRVI_DBG_FUSED
    const int dbg_tile = (int)((tile - tile_begin) >> 1);
#endif
    FT(0);
    // tiles_per_cell > 1 (S > 128, chain only: z_out, no distances): tile -> (cell, 128-row slice of its samples)
    const int64_t pos = tiles_per_cell > 1 ? tile / tiles_per_cell : tile * cells_per_tile + lc;  // sorted position of this row's cell
    const int sr = tiles_per_cell > 1 ? (int)(tile % tiles_per_cell) * kTcRows + r : s;
    const bool valid = tiles_per_cell > 1 ? (sr < S && pos < n_cells) : (row_in_tile && pos < n_cells);
    const int64_t cell = valid ? (gp.order ? (int64_t)gp.order[pos] : pos) : 0;
    const int ss = valid ? sr : 0;
    // sampled paths: every (cell, s) row carries its own draw of u (CellBuf::per_row); z_base then differs
    // between the rows of a cell and no longer cancels in the distances
    const int64_t fi = cb.per_row ? (valid ? cell * S + sr : 0) : cell;

    // ================= attention features of head `half` -> A1 chunks 2*half, 2*half+1 =================
    if (tc.att_tab != nullptr && !cb.per_row) {
      // table path (tc_build): the 16 features of this thread are 16 look-ups of ITS sample's table of cubics; the rows
      // of a cell share the cell's query tokens, so the lanes of a warp read consecutive 16-byte entries of one interval
      const float4* qp4 = reinterpret_cast<const float4*>(cb.q_tok + fi * 16);
      const float t0s = -tc.att_t0 * tc.att_scale;
      const float as = a_h * tc.att_scale, bs = fmaf(b_h, tc.att_scale, t0s);   // x = (alpha q + beta - t0) * scale
#pragma unroll 1
      for (int bt = 0; bt < 2; ++bt) {   // two batches of 8 features: 8 independent 16-byte loads in flight per thread
        const float4 qa = __ldg(qp4 + 2 * bt), qb = __ldg(qp4 + 2 * bt + 1);
        const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
        float4 nd[8];
        float fr[8];
#pragma unroll
        for (int f = 0; f < 8; ++f) {
          const float x = fminf(fmaxf(fmaf(as, qv[f], bs), 0.f), tc.att_xmax);
          const int k = (int)x;
          fr[f] = x - (float)k;
          nd[f] = __ldg(tc.att_tab + (uint32_t)(k * S + ss));   // uniform base + 32-bit index
        }
#pragma unroll
        for (int f2 = 0; f2 < 4; ++f2) {
          float mm[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float4 n4 = nd[2 * f2 + e];
            const float u = fr[2 * f2 + e];
            mm[e] = fmaf(u, fmaf(u, fmaf(u, n4.w, n4.z), n4.y), n4.x);   // entries are the cubic's coefficients
          }
          uint32_t hw, lw;
          split2(mm[0], mm[1], hw, lw);
          const int it = bt * 4 + f2;
          const int off = (half * 2 + (it >> 2)) * 2048 + (it & 3) * 4;
          *reinterpret_cast<uint32_t*>(row_a1h + off) = hw;
          *reinterpret_cast<uint32_t*>(row_a1l + off) = lw;
        }
      }
    } else {
      float2 kv2[8];  // kv * log2(e), pairs
      const float4* kp = reinterpret_cast<const float4*>(tc.kvl + (size_t)ss * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t4 = __ldg(kp + q);
        kv2[2 * q] = make_float2(t4.x, t4.y);
        kv2[2 * q + 1] = make_float2(t4.z, t4.w);
      }
      const float rmax = __ldg(tc.kvref + ss * 2), rmin = __ldg(tc.kvref + ss * 2 + 1);
      const float2* qp = reinterpret_cast<const float2*>(cb.q_tok + fi * 16);
      float2 qn = __ldg(qp);
#pragma unroll 1
      for (int it = 0; it < 8; ++it) {  // one packed word (two features) per iteration
        float mm[2];
        const float2 qc = qn;
        if (it < 7) qn = __ldg(qp + it + 1);   // the next pair's tokens travel while this pair's 32 exponentials run
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float t = fmaf(a_h, e ? qc.y : qc.x, b_h);
          const float off = -t * (t >= 0.f ? rmax : rmin);  // exponent <= 0 for every j
          const float2 t2 = make_float2(t, t), off2 = make_float2(off, off);
          float2 den = make_float2(0.f, 0.f), num = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 a = __ffma2_rn(t2, kv2[j], off2);
            const float2 ex = make_float2(ex2_approx(a.x), ex2_approx(a.y));
            den = __fadd2_rn(den, ex);
            num = __ffma2_rn(ex, kv2[j], num);
          }
          mm[e] = (num.x + num.y) * rcp_approx(den.x + den.y);  // = log2(e) * m[h,i]; 1/log2(e) is folded into B1 / B2a
        }
        uint32_t hw, lw;
        split2(mm[0], mm[1], hw, lw);
        const int off = (half * 2 + (it >> 2)) * 2048 + (it & 3) * 4;
        *reinterpret_cast<uint32_t*>(row_a1h + off) = hw;
        *reinterpret_cast<uint32_t*>(row_a1l + off) = lw;
      }
    }
    {
      if (half == 0) {  // constant-one column (k = 32) and zero padding of A1
        for (int c = 4; c < tc.K1 / 8; ++c)   // (the lo copy of A1 holds the feature chunks only: issue_gemm K_lo = 32)
          *reinterpret_cast<uint4*>(row_a1h + c * 2048) = make_uint4(c == 4 ? 0x00003C00u : 0u, 0u, 0u, 0u);
      }
    }
    FT(1);
    fence_proxy_async();
    wg_sync(wg);
    FT(2);

    // ================= GEMM1 / epilogue 1 =================
    if (issue_warp) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm(iu_tr0, iu_a1h, iu_a1l, s_w, tc.b1, tc.K1, 32, idesc128, true);
        umma_commit(iu_bar);
        issue_gemm_cat(iu_td2, iu_a1h, iu_a1l, s_w, tc.b2a, tc.K1, 32, idesc64, idesc32);
      }
      __syncwarp();
    }
    if (half == 1) {  // A3 static part under GEMM1 (GEMM3 is its first reader): column 32 = 1.0, columns 33.. = u_ln[0..Lq), then zeros
      const float* ul = cb.u_ln + fi * Lq;
      for (int c = 4; c < tc.K3 / 8; ++c) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
          const int k0 = (c - 4) * 8 + 2 * j2 - 1, k1 = k0 + 1;   // index into u_ln; -1 = the constant column
          const float a = k0 < 0 ? 1.0f : (k0 < Lq ? __ldg(ul + k0) : 0.f);
          const float b = k1 < Lq ? __ldg(ul + k1) : 0.f;
          split2(a, b, hw[j2], lw[j2]);
        }
        *reinterpret_cast<uint4*>(row_a3h + c * 2048) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(row_a3l + c * 2048) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    }
    FT(3);
    mbar_wait(bar_mma, phase); phase ^= 1;
    tc_fence_after();
    FT(4);
    ep128(t_lane, half, r, wg, g1, b1, xch);
    tc_fence_before();
    FT(5);
    wg_sync(wg);
    FT(6);
    // ================= GEMM2 / epilogue 2 =================
    if (issue_warp) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm_ts2_cat(iu_td2, iu_tr0, s_w, tc.b2t, idesc64, idesc32);
        umma_commit(iu_bar);
      }
      __syncwarp();
    }
    FT(7);
    mbar_wait(bar_mma, phase); phase ^= 1;
    tc_fence_after();
    FT(8);
    ep32x(t_d2_lane, half, r, wg, g2, b2, xch, row_a3h, row_a3l);
    tc_fence_before();
    fence_proxy_async();
    FT(9);
    wg_sync(wg);
    FT(10);
    // ================= GEMM3 / epilogue 3 =================
    if (issue_warp) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm(iu_tr0, iu_a3h, iu_a3l, s_w, tc.b3, tc.K3, tc.K3, idesc128, true);
        umma_commit(iu_bar);
        issue_gemm_cat(iu_td2, iu_a3h, iu_a3l, s_w, tc.b4a, tc.K3, tc.K3, idesc64, idesc32);
      }
      __syncwarp();
    }
    FT(11);
    mbar_wait(bar_mma, phase); phase ^= 1;
    tc_fence_after();
    FT(12);
    ep128(t_lane, half, r, wg, g3, b3, xch);
    tc_fence_before();
    FT(13);
    wg_sync(wg);
    FT(14);
    // ================= GEMM4 / epilogue 4 =================
    if (issue_warp) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm_ts2_cat(iu_td2, iu_tr0, s_w, tc.b4t, idesc64, idesc32);
        umma_commit(iu_bar);
      }
      __syncwarp();
    }
    FT(15);
    mbar_wait(bar_mma, phase); phase ^= 1;
    tc_fence_after();
    FT(16);
    // A5 = [t4 | 1 | 0] reuses the A1 buffers (GEMM2 was their last reader); the one / zero columns are still in place
    ep32x(t_d2_lane, half, r, wg, g4, b4, xch, row_a1h, row_a1l);
    tc_fence_before();
    fence_proxy_async();
    FT(17);
    wg_sync(wg);
    FT(18);
    // ================= GEMM5 / residual rows =================
    if (issue_warp) {
      tc_fence_after();
      if (elect_one()) {
        issue_gemm(iu_tr0, iu_a1h, iu_a1l, s_w, tc.b5, tc.K5, 32, idescN5, true);
        umma_commit(iu_bar);
      }
      __syncwarp();
    }
    FT(19);
    mbar_wait(bar_mma, phase); phase ^= 1;
    tc_fence_after();
    FT(20);
    // this thread's half of the N5 output columns stays in registers until the Gram operands are built
    const int ncol = tc.N5 >> 1, col0 = half * ncol;
    float v[32];
    if (ncol == 16) tmem_ld16(t_lane + col0, v); else tmem_ld32(t_lane + col0, v);
    tmem_ld_wait();
    // (columns >= L of D5 are exactly zero: the padded rows of the B5 operand are; only rows outside the chunk need zeroing)
    if (!valid) {
#pragma unroll
      for (int l = 0; l < 32; ++l) v[l] = 0.f;
    }
    if (cb.per_row && valid) {
      const float* zb = cb.zbase + fi * L;
#pragma unroll
      for (int l = 0; l < 32; ++l)
        if (l < ncol && col0 + l < L) v[l] += __ldg(zb + col0 + l);
    }
    tc_fence_before();
    if (z_out != nullptr) {
      // A1/A3 are dead (GEMM5 done): stage the fp32 residual tile on them, then write z = z_base + residual with
      // fully coalesced 4-byte stores: consecutive threads walk the tile row-major (a thread-per-row store would
      // touch 32 different rows, i.e. 32 sectors, per instruction)
#pragma unroll
      for (int l = 0; l < 32; ++l)
        if (l < ncol && col0 + l < Kz) zf[r * zld + col0 + l] = v[l];
      long long* rowinfo_w = reinterpret_cast<long long*>(xch);  // xch is free between the epilogues and the Gram
      if (half == 0) {  // row -> (output row, z_base row)
        rowinfo_w[r] = valid ? (long long)(cell * S + sr) : -1ll;
        rowinfo_w[128 + r] = (long long)fi;
      }
      wg_sync(wg);
      const long long* rowinfo = reinterpret_cast<const long long*>(xch);
      const bool add_zb = !cb.per_row;
      int r2 = wt / L, l2 = wt - r2 * L;
      const int dr = 256 / L, dl = 256 - dr * L;
      while (r2 < 128) {
        const long long orow = rowinfo[r2];
        if (orow >= 0) {
          float val = zf[r2 * zld + l2];
          if (add_zb) val += __ldg(cb.zbase + rowinfo[128 + r2] * L + l2);
          z_out[orow * L + l2] = val;
        }
        r2 += dr; l2 += dl;
        if (l2 >= L) { l2 -= L; ++r2; }
      }
    }
    if (!need_dist) {
      wg_sync(wg);  // D5 consumed, zf / rowinfo free before the next tile
      continue;
    }

    // ================= reference row per cell, norms, (hi, lo) Gram operands =================
    // Distances are invariant under a per-cell shift: every row is re-centred on the mean of two of the cell's rows (samples 0
    // and S / 2), which keeps the norms -- and with them the cancellation in n_i + n_j - 2 g -- small.  The closer the centre is
    // to the cell's mean, the fewer entries fall under the exact-recompute threshold (a ~1.5 K clk detour each, in groupby mode
    // under the accumulator lock): on random models with S = 128, 0.1-21 entries per cell with sample 0 alone, 0-3.7 with two
    // samples, 0-1.7 with the exact mean (which would cost a reduction over the rows).
    if (row_in_tile && (s == 0 || (kc2 && s == (S >> 1)))) {
      float* dst = cref_s + (lc * sm.cref_rows + (s == 0 ? 0 : 1)) * Kz + col0;   // [cell][cref_rows][Kz]
#pragma unroll
      for (int l = 0; l < 32; ++l)
        if (l < ncol) dst[l] = v[l];
    }
    FT(21);
    wg_sync(wg);  // reference rows visible; with z output: every thread is done with zf / rowinfo
    FT(22);
    {
      const float* crow = cref_s + (row_in_tile ? lc : 0) * sm.cref_rows * Kz + col0;
      const int c1off = kc2 ? Kz : 0;   // one reference row only (n_latent > 32: shared memory): the "second" row is the first
      const float2 mh2 = make_float2(-0.5f, -0.5f);
      float nrm_p = 0.f;
      const int nch = Kz >> 4;  // K chunks per half
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        if (cc < nch) {
          const int c = half * nch + cc;
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int j2 = 0; j2 < 4; ++j2) {
            const int k0 = cc * 8 + 2 * j2;
            const float2 cr0 = *reinterpret_cast<const float2*>(crow + k0), cr1 = *reinterpret_cast<const float2*>(crow + c1off + k0);
            const float2 ab = __ffma2_rn(cr1, mh2, __ffma2_rn(cr0, mh2, make_float2(v[k0], v[k0 + 1])));
            const float a = ab.x, b = ab.y;
            split2(a, b, hw[j2], lw[j2]);
            // norms from the values the tensor core sees (hi + lo), so that n_i + n_j - 2 g cancels consistently
            const float ar = h2f_lo(hw[j2]) + h2f_lo(lw[j2]), br = h2f_hi(hw[j2]) + h2f_hi(lw[j2]);
            nrm_p = fmaf(ar, ar, nrm_p);
            nrm_p = fmaf(br, br, nrm_p);
          }
          *reinterpret_cast<uint4*>(zh + c * 2048 + r * 16) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(zl + c * 2048 + r * 16) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
      xch[half * 128 + r].x = nrm_p;
    }
    FT(23);
    fence_proxy_async();
    wg_sync(wg);
    FT(24);
    // ================= Gram on the tensor cores =================
    if (issue_warp) {
      tc_fence_after();
      if (elect_one()) {
        // The two cross terms first, the Zh.Zh^T term last: G[i][j] and G[j][i] then differ only through the rounding of the
        // tiny (2^-11 relative) cross-term sum, i.e. by one ulp in rare cases, instead of through the order in which two
        // cross terms are added to a large accumulator in EVERY entry (near-symmetric output: D vs D^T)
        const uint32_t lbo = 128 * 16;
        for (int ks = 0; ks < Kz / 16; ++ks) {
          const uint64_t dh = umma_desc(iu_zh + ks * 2 * lbo, lbo, 128), dl = umma_desc(iu_zl + ks * 2 * lbo, lbo, 128);
          umma_f16(iu_tr0, dh, dl, idesc128, ks == 0 ? 0u : 1u);
          umma_f16(iu_tr0, dl, dh, idesc128, 1u);
        }
        for (int ks = 0; ks < Kz / 16; ++ks) {
          const uint64_t dh = umma_desc(iu_zh + ks * 2 * lbo, lbo, 128);
          umma_f16(iu_tr0, dh, dh, idesc128, 1u);
        }
        umma_commit(iu_bar);
      }
      __syncwarp();
    }
    // group-run bookkeeping, first half: the composite codes of the tile's first and last cell travel under the Gram MMAs.  The
    // cells are sorted by that code, so the tile lies inside one group run exactly when the two agree -- every thread sees it
    // without a reduction barrier.
    int first_code = -1, last_code = -1;
    if (gp.n_keys > 0) {
      first_code = gp.comp_sorted[min(tile * cells_per_tile, n_cells - 1)];
      last_code = gp.comp_sorted[min(tile * cells_per_tile + cells_per_tile - 1, n_cells - 1)];
    }
    const float n_i = xch[r].x + xch[128 + r].x;
    if (half == 0) nrm_s[r] = n_i;
    FT(25);
    mbar_wait(bar_mma, phase); phase ^= 1;
    tc_fence_after();
    FT(26);

    // ================= group-run bookkeeping =================
    bool use_acc = false;
    if (gp.n_keys > 0) {
      use_acc = first_code == last_code;
      if (!use_acc) wg_sync(wg);  // publishes nrm_s (the lock's barrier does otherwise)
      if (use_acc) {
        if (issuer) { while (atomicCAS(&ctrl->lock, 0, 1) != 0) __nanosleep(20); }
        wg_sync(wg);
        tc_fence_after();
        const int cur = *reinterpret_cast<volatile int*>(&ctrl->cur_code);
        const int cur_valid = *reinterpret_cast<volatile int*>(&ctrl->cur_valid);
        if (cur_valid && cur != first_code) {
          const long long rep = *reinterpret_cast<volatile long long*>(&ctrl->cur_cell);
          flush_acc(t_acc_lane, half, gp, rep, S, c0, row_in_tile, r);
        }
        wg_sync(wg);
        if (issuer) { ctrl->cur_code = first_code; ctrl->cur_valid = 1; ctrl->cur_cell = gp.order[tile * cells_per_tile]; }
      }
    } else {
      wg_sync(wg);  // publishes nrm_s
    }
    FT(27);
    // ================= distance epilogue: this thread's chunks of row r =================
#pragma unroll 1
    for (int c = ch_lo; c < ch_hi; ++c) {
      float g[32];
      tmem_ld32(t_lane + c * 32, g);
      tmem_ld_wait();
      // fast path (warp-uniform): the whole chunk lies inside every lane's cell and every row is valid -- no bounds checks;
      // the diagonal can only sit in the chunk of this warp's own rows (entry q == lane); cancellation-dominated entries
      // are only DETECTED here (min over the entries of d^2 - 2^-12 (n_i + n_j)) and located when there is one (rare)
      const bool full = __all_sync(0xffffffffu, valid && c * 32 >= c0 && c * 32 + 32 <= c0 + S);
      const float2 ni2 = make_float2(n_i, n_i), m2 = make_float2(-2.0f, -2.0f), thr2 = make_float2(-1.0f / 4096.0f, -1.0f / 4096.0f);
      if (full) {
        const float tmin = (c == warp_q) ? dist_chunk_fast<true>(g, nrm_s + c * 32, n_i, tid & 31)
                                         : dist_chunk_fast<false>(g, nrm_s + c * 32, n_i, tid & 31);
        if (__any_sync(0xffffffffu, tmin < 0.f)) {  // rare: near-duplicate samples -> exact direct-difference distance
          uint32_t flagged = 0u;
          if (tmin < 0.f) {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (c * 32 + q != r && g[q] * g[q] < (n_i + nrm_s[c * 32 + q]) * (1.0f / 4096.0f)) flagged |= 1u << q;
          }
          if (__any_sync(0xffffffffu, flagged != 0u)) exact_fix_warp<2048u>(g, flagged, zh, zl, r, c * 32, L);
        }
      } else {
        uint32_t flagged = 0u;
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 nj = *reinterpret_cast<const float4*>(nrm_s + c * 32 + 4 * q4);
          const float2 njp[2] = {make_float2(nj.x, nj.y), make_float2(nj.z, nj.w)};
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int q = 4 * q4 + 2 * e2, j = c * 32 + q;     // columns j, j+1 = tile rows of the other samples
            const float2 sn = __fadd2_rn(ni2, njp[e2]);
            const float2 d2 = __ffma2_rn(make_float2(g[q], g[q + 1]), m2, sn);
            const float2 tt = __ffma2_rn(sn, thr2, d2);        // < 0: cancellation dominated
            const bool in0 = valid && (unsigned)(j - c0) < (unsigned)S && j != r;
            const bool in1 = valid && (unsigned)(j + 1 - c0) < (unsigned)S && j + 1 != r;
            if (in0 && tt.x < 0.f) flagged |= 1u << q;
            if (in1 && tt.y < 0.f) flagged |= 2u << q;
            g[q] = in0 ? sqrt_approx(fmaxf(d2.x, 0.f)) : 0.f;
            g[q + 1] = in1 ? sqrt_approx(fmaxf(d2.y, 0.f)) : 0.f;
          }
        }
        if (__any_sync(0xffffffffu, flagged != 0u))  // rare: near-duplicate samples -> exact direct-difference distance
          exact_fix_warp<2048u>(g, flagged, zh, zl, r, c * 32, L);
      }
      if (cell_out != nullptr && (S & 3) == 0) {
        // Rows go through a per-warp shared-memory tile (the fp32 z tile is dead by now) so that every store
        // instruction writes full 32-byte sectors: a thread-per-row float4 store touches 32 rows = 32 half-used
        // sectors (measured on the isolated stage: 1.8 -> 2.3 TB/s written at S = 128).  Two 16-column sub-steps;
        // 16-byte units XOR-swizzled with the row: no bank conflicts on either side.  The rows of a warp may belong
        // to different cells: the destination row and the cell's first column travel by shuffle.
        const int lane = tid & 31;
        float4* T = reinterpret_cast<float4*>(scratch + sm.off_T) + (wt >> 5) * 128;
        const int my_row = valid ? (int)(cell * S + s) : -1;   // < 2^31: cells are indices inside one device chunk
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            T[lane * 4 + (q ^ ((lane >> 1) & 3))] = make_float4(g[h2 * 16 + 4 * q], g[h2 * 16 + 4 * q + 1], g[h2 * 16 + 4 * q + 2], g[h2 * 16 + 4 * q + 3]);
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rl = it * 8 + (lane >> 2), unit = lane & 3;
            const int orow = __shfl_sync(0xffffffffu, my_row, rl);
            const int oc0 = __shfl_sync(0xffffffffu, c0, rl);
            const int j = c * 32 + h2 * 16 + unit * 4 - oc0;
            const float4 v = T[rl * 4 + (unit ^ ((rl >> 1) & 3))];
            if (orow >= 0 && j >= 0 && j < S) *reinterpret_cast<float4*>(cell_out + (size_t)orow * S + j) = v;
          }
          __syncwarp();
        }
      } else if (cell_out != nullptr && valid) {
        float* dst = cell_out + ((size_t)cell * S + s) * S;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          const int j = c * 32 + q - c0;
          if (j >= 0 && j < S) dst[j] = g[q];
        }
      }
      if (use_acc) {
        float acc[32];
        tmem_ld32(t_acc_lane + c * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float2 t = __fadd2_rn(make_float2(acc[2 * q], acc[2 * q + 1]), make_float2(g[2 * q], g[2 * q + 1]));
          acc[2 * q] = t.x; acc[2 * q + 1] = t.y;
        }
        tmem_st32(t_acc_lane + c * 32, reinterpret_cast<uint32_t*>(acc));
      } else if (gp.n_keys > 0 && valid) {  // mixed tile: straight to the float64 sums
        for (int k = 0; k < gp.n_keys; ++k) {
          const int gk = gp.codes[k][cell];
          if (gk < 0 || gk >= gp.n_groups[k]) continue;
          double* base = gp.sums[k] + ((size_t)gk * S + s) * S;
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int j = c * 32 + q - c0;
            if (j >= 0 && j < S && j != s) atomicAdd(base + j, (double)g[q]);
          }
        }
      }
    }
    FT(28);
    if (use_acc) {
      tmem_st_wait();
      tc_fence_before();
      wg_sync(wg);
      if (issuer) { __threadfence_block(); atomicExch(&ctrl->lock, 0); }
    }
    tc_fence_before();
    wg_sync(wg);  // Gram / scratch fully consumed before the next tile reuses them
    FT(29);
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (gp.n_keys > 0 && wg == 0 && ctrl->cur_valid) {
    flush_acc(t_acc_lane, half, gp, ctrl->cur_cell, S, c0, row_in_tile, r);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

inline bool tc_fused_supported(const TcNet& tc, int S, int norm) { return tc.ready && norm == MRVI_NORM_L2 && S >= 4 && S <= 128; }

inline int tc_fused_prepare(TcNet& tc, int L) {
  const size_t bytes = fused_smem_map(tc.wimg_bytes, tc.K1, tc.K3, L).total + 1024;
  if (bytes > 227 * 1024) return 0;  // fused path unavailable for this shape; the unfused kernels are used
  if (cudaFuncSetAttribute(k_chain_fused_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin) != cudaSuccess) {
    g_tc_error = "cudaFuncSetAttribute(k_chain_fused_tc) failed";
    return MRVI_ERR_CUDA;
  }
  tc.fused_smem_bytes = bytes;
  return 0;
}

inline void tc_launch_fused(const TcNet& tc, const NetW& net, int64_t n, const CellBuf& cb, float* z_out, float* cell_out,
                            const GroupPlan& gp, cudaStream_t st) {
  const int S = net.S;
  const int cpt = kTcRows / S;
  const int64_t n_tiles = (n + cpt - 1) / cpt;
  if (n_tiles == 0) return;
  const int grid = (int)std::min<int64_t>((n_tiles + 1) / 2, tc.sm_count);
  const int64_t tiles_per_cta = (n_tiles + grid - 1) / grid;
  k_chain_fused_tc<<<grid, kFusedThreads, tc.fused_smem_bytes, st>>>(tc, cb, n, S, net.L, net.Lq, z_out, cell_out, gp,
                                                                    n_tiles, cpt, tiles_per_cta, 1);
}

// chain only (z_out, no distances) through the same kernel, any S: for S > 128 a cell spans several 128-row tiles
inline void tc_launch_fused_chain(const TcNet& tc, const NetW& net, int64_t n, const CellBuf& cb, float* z_out, cudaStream_t st) {
  const int S = net.S;
  int cpt = 1, tpc = 1;
  int64_t n_tiles;
  if (S <= kTcRows) { cpt = kTcRows / S; n_tiles = (n + cpt - 1) / cpt; }
  else { tpc = (S + kTcRows - 1) / kTcRows; n_tiles = n * tpc; }
  if (n_tiles == 0) return;
  const int grid = (int)std::min<int64_t>((n_tiles + 1) / 2, tc.sm_count);
  const int64_t tiles_per_cta = (n_tiles + grid - 1) / grid;
  GroupPlan none;
  memset(&none, 0, sizeof(none));
  k_chain_fused_tc<<<grid, kFusedThreads, tc.fused_smem_bytes, st>>>(tc, cb, n, S, net.L, net.Lq, z_out, nullptr, none, n_tiles,
                                                                    cpt, tiles_per_cta, tpc);
}

}  // namespace mrvi
