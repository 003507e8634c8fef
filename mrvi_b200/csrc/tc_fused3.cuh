// Fused tensor-core kernel, third version: q(z|u,s) chain + pairwise distances + group reduction with WARP-SPECIALISED
// roles (round 2).  Same arithmetic as k_chain_fused_tc (tc_fused.cuh); what changed is who does what, driven by
// tools/mma_microbench.cu and the round-1 ncu profile (no pipe above 46 % busy: latency- and sync-bound):
//
//   * MMA issue.  A `tcgen05.mma` issued from `if (thread == 0)` code costs ~92 clk per instruction: the compiler
//     cannot prove the descriptors warp-uniform and wraps every UTCHMMA in an ELECT / R2UR waterfall loop.  Issued from
//     a warp-uniform branch (warp index through a shuffle, `elect.sync` around the instruction only) the same train
//     runs at 42 clk per MMA for N <= 64 and 61 clk for N = 128 (= the tensor-pipe floor).  Warp 16 is a dedicated
//     issuer for both tile slots: it polls the slots' "operands ready" mbarriers, issues the next GEMM of whichever
//     slot is ready and commits to that slot's "accumulator ready" mbarrier.  The epilogue warps never issue, and the
//     barrier in front of every GEMM becomes a non-blocking arrive.
//   * Fewer, wider MMAs.  The 32-wide operands are stored as [W_hi | W_lo] concatenated along N (tc_kernels.cuh TcOp),
//     so A.W_hi and A.W_lo come out of ONE N = 64 instruction into columns [0,32) / [32,64) of the D2 region (the
//     32-wide epilogue adds the two halves): GEMM2 / GEMM4 take 16 + 5 instead of 24 + 8 instructions.
//   * Early issue.  The skip-connection parts of GEMM2 / GEMM4 (A1.B2a, A3.B4a) do not depend on the epilogue in
//     front of them: they are issued right behind GEMM1 / GEMM3 and run under epilogue 1 / 3.
//   * Attention off the critical path.  The 32 softmax-weighted means per row (512 MUFU.EX2 per row, 25 % of the
//     round-1 tile time, executed serially ahead of GEMM1 by the same warps) are produced by warps 17-19 for the NEXT
//     tile of a slot while its epilogue warpgroup works on the current one (A1 is free again as soon as GEMM2 has
//     read it: A5 = [t4 | 1] now reuses the A3 buffer, whose constant column sits at the same K index).
//   * Registers follow the roles (`setmaxnreg`): 104 for the 16 epilogue warps, 64 for the issuer / producer warpgroup.
//
// Block = 640 threads: warps 0-7 / 8-15 = epilogue warpgroups of tile slots 0 / 1 (two threads per row, as before),
// warp 16 = issuer, warps 17-19 = attention producers.  TMEM: slot s: [192 s, +128) D1/T/D3/T/D5/Gram, [192 s + 128,
// +64) D2/D4 as [A.W_hi | A.W_lo]; [384, 512) group accumulator.
// Supported: everything the second version supports except fused distances with n_latent > 32 (tc_fused.cuh keeps
// those); chain-only mode (z output, any S) writes z straight from registers.
#pragma once
#include "tc_encoder.cuh"   // tmem_st16
#include "tc_fused.cuh"

namespace mrvi {

#ifdef MRVI_DBG_F3
// clock64 timeline of CTA 0 (tools/dbg/run_dbg_f3.py): [wg][tile < 16][slot < 32] epilogue stamps, then issuer / producer stamps
__device__ long long g_dbg_f3[2 * 16 * 32 + 2 * 16 * 16 + 4 * 64 * 2];
#define F3_STAMP(wg_, k_, slot_) do { if (blockIdx.x == 0 && (k_) < 16 && (tid & 255) == 0) g_dbg_f3[((wg_) * 16 + (k_)) * 32 + (slot_)] = clock64(); } while (0)
#define F3_ISTAMP(s_, k_, step_, which_) do { if (blockIdx.x == 0 && (k_) < 16 && lane == 0) g_dbg_f3[1024 + ((s_) * 16 + (k_)) * 16 + (step_) * 2 + (which_)] = clock64(); } while (0)
#define F3_PSTAMP(pw_, t_, which_) do { if (blockIdx.x == 0 && (t_) < 64 && lane == 0) g_dbg_f3[1536 + ((pw_) * 64 + (t_)) * 2 + (which_)] = clock64(); } while (0)
#else
#define F3_STAMP(wg_, k_, slot_) do { } while (0)
#define F3_ISTAMP(s_, k_, step_, which_) do { } while (0)
#define F3_PSTAMP(pw_, t_, which_) do { } while (0)
#endif

constexpr int kF3Threads = 640;       // warps 0-15 epilogue, 16 MMA issuer, 17-19 attention producers
constexpr int kF3ProducerWarps = 3;
#ifndef F3_EPI_PAIRS
#define F3_EPI_PAIRS 0   // measured (C2 / C4 shard / C3 shard fused stage, ms): 0 -> 1.81 / 1.76 / 1.60, 4 -> 2.02 / 1.98 / 1.94
#endif
#ifndef F3_LN_BARRIER
#define F3_LN_BARRIER 1   // statistics exchanged behind a barrier: 1.81 / 1.76 / 1.60; both threads of a row read the whole row instead: 2.02 / 1.90 / 1.60
#endif
constexpr int kF3EpiPairs = F3_EPI_PAIRS;        // attention feature words 0..3 of a row: its epilogue threads, inside their MMA waits; 4..7: the producer warps
// setmaxnreg only redistributes what the launch allocated (640 threads x 96 registers = 61440): 512 x 104 + 128 x 64 =
// 61440.  (A 21st warp would drop the launch allocation to 80 registers per thread -- warps are allocated in fours --
// and leave the epilogue warps short: measured with -Xptxas -v.)
constexpr int kF3RegsEpi = 104, kF3RegsAux = 64;

struct F3Smem {
  uint32_t w;                            // weight image
  uint32_t slot[2];                      // per tile slot: [a1h | a1l (4 feature chunks) | a3h | a3l]
  uint32_t off_a1l, off_a3h, off_a3l;    // relative to the slot base
  uint32_t off_zh, off_zl, off_T, off_cref;  // Gram operands, keep_cell staging tiles and reference rows: over a3h / a3l (dead after GEMM5)
  uint32_t nrm[2], xch[2], lnv, ctrl, total;
};

__host__ __device__ inline F3Smem f3_smem_map(uint32_t wimg_bytes, int K1, int K3) {
  F3Smem m;
  uint32_t o = 0;
  auto up = [](uint32_t b) { return (b + 1023u) & ~1023u; };
  m.w = o; o += up(wimg_bytes);
  const uint32_t a1 = 128u * K1 * 2u, a1l = 128u * 32u * 2u, a3 = 128u * K3 * 2u;
  m.off_a1l = a1; m.off_a3h = a1 + a1l; m.off_a3l = a1 + a1l + a3;
  m.off_zh = m.off_a3h; m.off_zl = m.off_zh + 8192u; m.off_T = m.off_zl + 8192u; m.off_cref = m.off_T + 8192u;
  uint32_t slot_bytes = m.off_a3l + a3;
  if (slot_bytes < m.off_cref + 4096u) slot_bytes = m.off_cref + 4096u;
  m.slot[0] = o; o += slot_bytes;
  m.slot[1] = o; o += slot_bytes;
  m.nrm[0] = o; o += 512;
  m.nrm[1] = o; o += 512;
  m.xch[0] = o; o += 2048;
  m.xch[1] = o; o += 2048;
  m.lnv = o; o += (uint32_t)((sizeof(TcLn) + 127) & ~size_t(127));
  m.ctrl = o; o += 256;
  m.total = o;
  return m;
}

struct F3Ctrl {
  uint64_t bar_w;
  uint64_t mma_done[2];   // issuer's tcgen05.commit -> epilogue warpgroup of the slot
  uint64_t ep_done[2];    // epilogue warpgroup (8 warp arrivals) -> issuer: operands of the slot's next GEMM are in place
  uint64_t a1_full[2];    // producers (8 task arrivals per tile) -> issuer: attention features of the slot's next tile
  uint64_t a1_empty[2];   // issuer's commit behind GEMM2 -> producers: A1 of the slot may be overwritten
  uint32_t tmem_slot;
  int lock;
  int cur_code;
  int cur_valid;
  long long cur_cell;
};

// one non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"   // test_wait, NOT try_wait: try_wait may suspend the
      "selp.b32 %0, 1, 0, p;\n\t}"                                   // warp for a time-out while the OTHER slot is ready
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
template <int R>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

// Partial LayerNorm sums (sum, sum of squares) of 32 consecutive accumulator columns of this thread's lane.
__device__ __forceinline__ float2 ln_partial32(uint32_t taddr) {
  float v[32];
  tmem_ld32(taddr, v);
  tmem_ld_wait();
  float2 a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = make_float2(v[2 * i], v[2 * i + 1]);
    a1 = __fadd2_rn(a1, x);
    a2 = __ffma2_rn(x, x, a2);
  }
  return make_float2(a1.x + a1.y, a2.x + a2.y);
}

// LayerNorm + gelu of the 128-column accumulator row, this thread's 64-column half, written back in place per 32-column
// chunk as [16 hi words | 16 lo words].  Both threads of a row compute the row statistics over ALL 128 columns
// themselves -- per-chunk partial sums with one instruction sequence, combined as (own half) + (other half), which is
// commutative, so the two threads get bit-identical statistics -- instead of exchanging partial sums through shared
// memory behind a 256-thread barrier in the middle of every epilogue (round 1: 4 such barriers per tile; stall_barrier
// ~1 per issue).  Cost: the statistics pass reads the row twice (+64 KB of TMEM reads per tile at ~800 B/clk).
__device__ __forceinline__ void ep128r(uint32_t t_lane, int half, const float* g_s, const float* b_s) {
  const float2 o0 = ln_partial32(t_lane + (2 * half) * 32), o1 = ln_partial32(t_lane + (2 * half + 1) * 32);
  const float2 p0 = ln_partial32(t_lane + (2 * (half ^ 1)) * 32), p1 = ln_partial32(t_lane + (2 * (half ^ 1) + 1) * 32);
  const float s1 = (o0.x + o1.x) + (p0.x + p1.x), s2 = (o0.y + o1.y) + (p0.y + p1.y);
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

// 16 columns of the 32-wide accumulator stored as two halves [A.W_hi | A.W_lo] (columns [0,32) / [32,64)): x = sum of
// the halves; returns the partial LayerNorm sums
__device__ __forceinline__ float2 ln_partial16x(uint32_t t_d2_lane, int col, float (&x)[16]) {
  float w[16];
  tmem_ld16(t_d2_lane + col, x);
  tmem_ld16(t_d2_lane + 32 + col, w);
  tmem_ld_wait();
  float2 a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 t = __fadd2_rn(make_float2(x[2 * i], x[2 * i + 1]), make_float2(w[2 * i], w[2 * i + 1]));
    x[2 * i] = t.x; x[2 * i + 1] = t.y;
    a1 = __fadd2_rn(a1, t);
    a2 = __ffma2_rn(t, t, a2);
  }
  return make_float2(a1.x + a1.y, a2.x + a2.y);
}

// LayerNorm + gelu of that 32-wide accumulator: this thread's 16 columns -> two 16-byte K chunks (hi and lo) of the
// next A operand in shared memory.  Row statistics as in ep128r (no exchange, no barrier).
__device__ __forceinline__ void ep32xr(uint32_t t_d2_lane, int half, const float* g_s, const float* b_s, uint8_t* row_h, uint8_t* row_l) {
  float v[16], other[16];
  const float2 po = ln_partial16x(t_d2_lane, (half ^ 1) * 16, other);
  const float2 pm = ln_partial16x(t_d2_lane, half * 16, v);
  const float s1 = pm.x + po.x, s2 = pm.y + po.y;
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

// One packed word of attention features (two softmax-weighted means of the per-sample kv tokens, head of `a_h / b_h`)
// -> hi / lo words of A1.  kv2: kv * log2(e); the result is log2(e) * m (1 / log2(e) is folded into B1 / B2a).
__device__ __forceinline__ void f3_feature_pair(const float2 (&kv2)[8], float q0, float q1, float a_h, float b_h, float rmax, float rmin,
                                                uint32_t* dst_h, uint32_t* dst_l) {
  float mm[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const float t = fmaf(a_h, e ? q1 : q0, b_h);
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
    mm[e] = (num.x + num.y) * rcp_approx(den.x + den.y);
  }
  uint32_t hw, lw;
  split2(mm[0], mm[1], hw, lw);
  *dst_h = hw;
  *dst_l = lw;
}

// ---- MMA trains of the issuer (called by one elected lane of a warp-uniform branch) ----
// Descriptors of consecutive K slices differ by a constant in the 14-bit start-address field (shared memory is below
// 256 KB: no carry): one base descriptor per operand, one 32-bit add per slice -- the issuer shares its scheduler with
// four epilogue warps, every instruction it saves is issue latency.
__device__ __forceinline__ uint64_t desc_adv(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// [128 x K] (hi, lo in shared memory) x [K x 128] (hi, lo): 3 terms per K slice
__device__ __forceinline__ void f3_gemm_ss128(uint32_t d, uint32_t a_h, uint32_t a_l, uint32_t s_w, const TcOp& b, int K, int K_lo,
                                              uint32_t idesc) {
  uint64_t bd = umma_desc(s_w + b.h, b.lbo, 128), bl = umma_desc(s_w + b.l, b.lbo, 128);
  uint64_t ah = umma_desc(a_h, 2048, 128), al = umma_desc(a_l, 2048, 128);
  for (int ks = 0; ks < K / 16; ++ks) {
    umma_f16(d, ah, bd, idesc, ks == 0 ? 0u : 1u);
    if (ks * 16 < K_lo) umma_f16(d, al, bd, idesc, 1u);
    umma_f16(d, ah, bl, idesc, 1u);
    bd = desc_adv(bd, 2 * b.lbo); bl = desc_adv(bl, 2 * b.lbo); ah = desc_adv(ah, 4096); al = desc_adv(al, 4096);
  }
}
// [128 x K] (hi, lo in shared memory) x concatenated [K x (32 | 32)]: A_hi.[W_hi | W_lo] (N = 64) + A_lo.W_hi (N = 32); clears D
__device__ __forceinline__ void f3_gemm_ss_cat(uint32_t d, uint32_t a_h, uint32_t a_l, uint32_t s_w, const TcOp& b, int K, int K_lo,
                                               uint32_t idesc64, uint32_t idesc32) {
  uint64_t bd = umma_desc(s_w + b.h, b.lbo, 128);
  uint64_t ah = umma_desc(a_h, 2048, 128), al = umma_desc(a_l, 2048, 128);
  for (int ks = 0; ks < K / 16; ++ks) {
    umma_f16(d, ah, bd, idesc64, ks == 0 ? 0u : 1u);
    if (ks * 16 < K_lo) umma_f16(d, al, bd, idesc32, 1u);
    bd = desc_adv(bd, 2 * b.lbo); ah = desc_adv(ah, 4096); al = desc_adv(al, 4096);
  }
}
// [128 x 128] (A in TMEM, chunked [hi16 | lo16] layout) x concatenated [128 x (32 | 32)], accumulating
__device__ __forceinline__ void f3_gemm_ts_cat(uint32_t d, uint32_t a_tmem, uint32_t s_w, const TcOp& b, uint32_t idesc64, uint32_t idesc32) {
  uint64_t bd = umma_desc(s_w + b.h, b.lbo, 128);
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint32_t a = a_tmem + (ks >> 1) * 32 + (ks & 1) * 8;
    umma_f16_ts(d, a, bd, idesc64, 1u);
    umma_f16_ts(d, a + 16, bd, idesc32, 1u);
    bd = desc_adv(bd, 2 * b.lbo);
  }
}

__global__ void __launch_bounds__(kF3Threads, 1)
k_chain_fused3_tc(const __grid_constant__ TcNet tc, CellBuf cb, int64_t n_cells, int S, int L, int Lq,
                  float* __restrict__ z_out, float* __restrict__ cell_out, GroupPlan gp, int64_t n_tiles,
                  int cells_per_tile, int64_t tiles_per_cta, int tiles_per_cell) {
  extern __shared__ __align__(1024) uint8_t smem_tc[];
  uint8_t* const smem = smem_tc;
  const F3Smem sm = f3_smem_map(tc.wimg_bytes, tc.K1, tc.K3);
  F3Ctrl* ctrl = reinterpret_cast<F3Ctrl*>(smem + sm.ctrl);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
  const int lane = tid & 31;
  float* lnv = reinterpret_cast<float*>(smem + sm.lnv);
  const bool need_dist = (cell_out != nullptr) || (gp.n_keys > 0);

  if (tid == 0) {
    mbar_init(&ctrl->bar_w, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctrl->mma_done[s], 1);
      mbar_init(&ctrl->ep_done[s], 8);
      mbar_init(&ctrl->a1_full[s], 8);
      mbar_init(&ctrl->a1_empty[s], 1);
    }
    ctrl->lock = 0; ctrl->cur_code = -1; ctrl->cur_valid = 0; ctrl->cur_cell = 0;
    fence_mbar_init();
  }
  for (int i = tid; i < (int)(sizeof(TcLn) / sizeof(float)); i += kF3Threads) lnv[i] = reinterpret_cast<const float*>(&tc.ln)[i];
  if (tid < 256) {  // constant-one column (k = 32) and zero padding of A1 (hi copy; the lo copy has the feature chunks only): once
    uint8_t* a1h = smem + sm.slot[tid >> 7] + (tid & 127) * 16;
    for (int c = 4; c < tc.K1 / 8; ++c) *reinterpret_cast<uint4*>(a1h + c * 2048) = make_uint4(c == 4 ? 0x00003C00u : 0u, 0u, 0u, 0u);
  }
  if (warp == 0) tmem_alloc(&ctrl->tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(&ctrl->bar_w, tc.wimg_bytes);
    bulk_g2s(smem + sm.w, tc.wimg, tc.wimg_bytes, &ctrl->bar_w);
  }
  if (gp.n_keys > 0 && warp < 8) {  // clear the group accumulator once
    const uint32_t t_acc = tmem_base + 384 + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t z[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) z[q] = 0u;
    tmem_st32(t_acc + (2 * (warp >> 2)) * 32, z);
    tmem_st32(t_acc + (2 * (warp >> 2) + 1) * 32, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int64_t tile_begin = (int64_t)blockIdx.x * tiles_per_cta;
  const int64_t tile_end = min(n_tiles, tile_begin + tiles_per_cta);
  const int64_t my_tiles = tile_end > tile_begin ? tile_end - tile_begin : 0;

  if (warp >= 16) {
    setmaxnreg_dec<kF3RegsAux>();
    if (warp == 16) {
      // ============================================ MMA issuer ============================================
      const uint32_t s_w = smem_u32(smem + sm.w);
      const uint32_t idesc128 = umma_idesc_f16(128), idesc64 = umma_idesc_f16(64), idesc32 = umma_idesc_f16(32),
                     idescN5 = umma_idesc_f16(tc.N5);
      const int nsteps = need_dist ? 6 : 5;
      int64_t left0 = (my_tiles + 1) >> 1, left1 = my_tiles >> 1;
      int step0 = 0, step1 = 0;
      uint32_t ph_ep0 = 0, ph_ep1 = 0, ph_a0 = 0, ph_a1 = 0;
      mbar_wait(&ctrl->bar_w, 0);
      while (left0 > 0 || left1 > 0) {
        bool any = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          int64_t& left = s ? left1 : left0;
          int& step = s ? step1 : step0;
          uint32_t& ph_ep = s ? ph_ep1 : ph_ep0;
          uint32_t& ph_a = s ? ph_a1 : ph_a0;
          if (left <= 0) continue;
          // (probe results go through a shuffle so that the compiler sees a warp-uniform branch: no ELECT waterfall
          //  around the UTCHMMA instructions below)
          if (!__shfl_sync(0xffffffffu, (int)mbar_test(&ctrl->ep_done[s], ph_ep), 0)) continue;
          if (step == 0 && !__shfl_sync(0xffffffffu, (int)mbar_test(&ctrl->a1_full[s], ph_a), 0)) continue;
          tc_fence_after();
          F3_ISTAMP(s, ((s ? (my_tiles >> 1) : ((my_tiles + 1) >> 1)) - left), step, 0);
          const uint32_t t_d1 = __shfl_sync(0xffffffffu, tmem_base, 0) + s * 192, t_d2 = t_d1 + 128;
          const uint32_t sa1h = smem_u32(smem + sm.slot[s]), sa1l = sa1h + sm.off_a1l, sa3h = sa1h + sm.off_a3h, sa3l = sa1h + sm.off_a3l;
          if (elect_one()) {
            if (step == 0) {
              f3_gemm_ss128(t_d1, sa1h, sa1l, s_w, tc.b1, tc.K1, 32, idesc128);
              umma_commit(&ctrl->mma_done[s]);
              f3_gemm_ss_cat(t_d2, sa1h, sa1l, s_w, tc.b2a, tc.K1, 32, idesc64, idesc32);   // skip part of GEMM2, under epilogue 1
            } else if (step == 1) {
              f3_gemm_ts_cat(t_d2, t_d1, s_w, tc.b2t, idesc64, idesc32);
              umma_commit(&ctrl->mma_done[s]);
              umma_commit(&ctrl->a1_empty[s]);                                               // A1 has been read for the last time
            } else if (step == 2) {
              f3_gemm_ss128(t_d1, sa3h, sa3l, s_w, tc.b3, tc.K3, tc.K3, idesc128);
              umma_commit(&ctrl->mma_done[s]);
              f3_gemm_ss_cat(t_d2, sa3h, sa3l, s_w, tc.b4a, tc.K3, tc.K3, idesc64, idesc32);  // skip part of GEMM4, under epilogue 3
            } else if (step == 3) {
              f3_gemm_ts_cat(t_d2, t_d1, s_w, tc.b4t, idesc64, idesc32);
              umma_commit(&ctrl->mma_done[s]);
            } else if (step == 4) {
              f3_gemm_ss128(t_d1, sa3h, sa3l, s_w, tc.b5, tc.K5, 32, idescN5);                // A5 = [t4 | 1 | ...] lives in the A3 buffer
              umma_commit(&ctrl->mma_done[s]);
            } else {
              const uint32_t zh = sa1h + sm.off_zh, zl = sa1h + sm.off_zl;
              uint64_t dh = umma_desc(zh, 2048, 128), dl = umma_desc(zl, 2048, 128);
              const uint64_t dh0 = dh;
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {   // cross terms first, Zh.Zh^T last (see tc_fused.cuh: near-symmetric output)
                umma_f16(t_d1, dh, dl, idesc128, ks == 0 ? 0u : 1u);
                umma_f16(t_d1, dl, dh, idesc128, 1u);
                dh = desc_adv(dh, 4096); dl = desc_adv(dl, 4096);
              }
              umma_f16(t_d1, dh0, dh0, idesc128, 1u);
              umma_f16(t_d1, desc_adv(dh0, 4096), desc_adv(dh0, 4096), idesc128, 1u);
              umma_commit(&ctrl->mma_done[s]);
            }
          }
          __syncwarp();
          F3_ISTAMP(s, ((s ? (my_tiles >> 1) : ((my_tiles + 1) >> 1)) - left), step, 1);
          ph_ep ^= 1;
          if (step == 0) ph_a ^= 1;
          if (++step == nsteps) { step = 0; --left; }
          any = true;
        }
        if (!any) __nanosleep(32);   // nothing ready: leave the issue slots of this scheduler to the epilogue warps
      }
    } else {
      // ======================================== attention producers ========================================
      // task = (tile of this CTA, 32-row group, head): 16 softmax-weighted means per row -> A1 chunks 2 head, 2 head + 1
      const int pw = warp - 17;
      for (int64_t task = pw; task < my_tiles * 8; task += kF3ProducerWarps) {
        const int64_t tseq = task >> 3;
        const int sub = (int)(task & 7), s = (int)(tseq & 1), head = sub & 1;
        const int64_t k = tseq >> 1, tile = tile_begin + tseq;
        const int r = (sub >> 1) * 32 + lane;
        // row -> (cell, sample): as the epilogue warpgroups
        const int lc = r / S, sl = r - lc * S;
        const int64_t pos = tiles_per_cell > 1 ? tile / tiles_per_cell : tile * cells_per_tile + lc;
        const int sr = tiles_per_cell > 1 ? (int)(tile % tiles_per_cell) * kTcRows + r : sl;
        const bool valid = tiles_per_cell > 1 ? (sr < S && pos < n_cells) : (lc < cells_per_tile && pos < n_cells);
        const int64_t cell = valid ? (gp.order ? (int64_t)gp.order[pos] : pos) : 0;
        const int ss = valid ? sr : 0;
        const int64_t fi = cb.per_row ? (valid ? cell * S + sr : 0) : cell;
        float2 kv2[8];
        float qv[16 - 2 * kF3EpiPairs];
        {
          const float4* kp = reinterpret_cast<const float4*>(tc.kvl + (size_t)ss * 16);
          const float2* qp = reinterpret_cast<const float2*>(cb.q_tok + fi * 16 + 2 * kF3EpiPairs);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t4 = __ldg(kp + q);
            kv2[2 * q] = make_float2(t4.x, t4.y);
            kv2[2 * q + 1] = make_float2(t4.z, t4.w);
          }
#pragma unroll
          for (int q = 0; q < 8 - kF3EpiPairs; ++q) {
            const float2 q2 = __ldg(qp + q);
            qv[2 * q] = q2.x; qv[2 * q + 1] = q2.y;
          }
        }
        const float rmax = __ldg(tc.kvref + ss * 2), rmin = __ldg(tc.kvref + ss * 2 + 1);
        const float a_h = head ? tc.alpha[1] : tc.alpha[0], b_h = head ? tc.beta[1] : tc.beta[0];
        mbar_wait(&ctrl->a1_empty[s], (uint32_t)((k & 1) ^ 1));   // GEMM2 of the slot's previous tile has read A1
        F3_PSTAMP(pw, task / kF3ProducerWarps, 0);
        uint8_t* row_h = smem + sm.slot[s] + r * 16;
        uint8_t* row_l = row_h + sm.off_a1l;
#pragma unroll
        for (int it = kF3EpiPairs; it < 8; ++it) {  // one packed word (two features) per iteration
          const int off = (head * 2 + (it >> 2)) * 2048 + (it & 3) * 4;
          f3_feature_pair(kv2, qv[2 * (it - kF3EpiPairs)], qv[2 * (it - kF3EpiPairs) + 1], a_h, b_h, rmax, rmin,
                          reinterpret_cast<uint32_t*>(row_h + off), reinterpret_cast<uint32_t*>(row_l + off));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctrl->a1_full[s]);
        F3_PSTAMP(pw, task / kF3ProducerWarps, 1);
      }
    }
  } else {
    // ============================================ epilogue warpgroups ============================================
    setmaxnreg_inc<kF3RegsEpi>();
    const int wg = warp >> 3, wt = tid & 255, half = wt >> 7, r = wt & 127, warp_q = (wt >> 5) & 3;
    const float *g1 = lnv, *b1 = lnv + 128, *g2 = lnv + 256, *b2 = lnv + 288, *g3 = lnv + 320, *b3 = lnv + 448,
                *g4 = lnv + 576, *b4 = lnv + 608;
    const uint32_t t_r0 = tmem_base + wg * 192, t_d2 = t_r0 + 128;
    const uint32_t lane_off = (uint32_t)(warp_q * 32) << 16;
    const uint32_t t_lane = t_r0 + lane_off, t_d2_lane = t_d2 + lane_off, t_acc_lane = tmem_base + 384 + lane_off;
    uint8_t* slot = smem + sm.slot[wg];
    uint8_t* row_a3h = slot + sm.off_a3h + r * 16;
    uint8_t* row_a3l = slot + sm.off_a3l + r * 16;
    uint8_t* zh = slot + sm.off_zh;
    uint8_t* zl = slot + sm.off_zl;
    float* cref_s = reinterpret_cast<float*>(slot + sm.off_cref);
    float* nrm_s = reinterpret_cast<float*>(smem + sm.nrm[wg]);
    float2* xch = reinterpret_cast<float2*>(smem + sm.xch[wg]);
    uint64_t* bar_mma = &ctrl->mma_done[wg];
    uint64_t* bar_ep = &ctrl->ep_done[wg];
    uint32_t phase = 0;
    const bool leader = wt == 0;   // lock / run-state bookkeeping of the warpgroup
    auto ep_arrive = [&]() {       // "this warp's part of the next GEMM's operands is in place" (8 warp arrivals per phase)
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ep);
    };

    // static row -> (local cell, sample) map of this thread
    const int lc = r / S, s = r - lc * S, c0 = lc * S;
    const bool row_in_tile = lc < cells_per_tile;
    uint8_t* row_a1h = slot + r * 16;
    uint8_t* row_a1l = slot + sm.off_a1l + r * 16;
    const float att_a = half ? tc.alpha[1] : tc.alpha[0], att_b = half ? tc.beta[1] : tc.beta[0];
    int pairs_next = 0;   // feature words 0 .. kF3EpiPairs-1 of this thread's row already written for the slot's NEXT tile
    // This thread's share of the attention features of tile `tn` (head = half), one packed word at a time; with `bar`
    // it stops as soon as that mbarrier phase completes (the work fills an MMA wait and must not delay the epilogue
    // behind it).  Warp-uniform: pairs_next and the barrier probe are the same for all lanes.
    auto attn_fill = [&](int64_t tn, uint64_t* bar, uint32_t ph) {
      if (pairs_next >= kF3EpiPairs || tn >= tile_end) return;
      const int64_t pos_n = tiles_per_cell > 1 ? tn / tiles_per_cell : tn * cells_per_tile + lc;
      const int sr_n = tiles_per_cell > 1 ? (int)(tn % tiles_per_cell) * kTcRows + r : s;
      const bool valid_n = tiles_per_cell > 1 ? (sr_n < S && pos_n < n_cells) : (row_in_tile && pos_n < n_cells);
      const int64_t cell_n = valid_n ? (gp.order ? (int64_t)gp.order[pos_n] : pos_n) : 0;
      const int ss_n = valid_n ? sr_n : 0;
      const int64_t fi_n = cb.per_row ? (valid_n ? cell_n * S + sr_n : 0) : cell_n;
      float2 kv2[8];
      const float4* kp = reinterpret_cast<const float4*>(tc.kvl + (size_t)ss_n * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t4 = __ldg(kp + q);
        kv2[2 * q] = make_float2(t4.x, t4.y);
        kv2[2 * q + 1] = make_float2(t4.z, t4.w);
      }
      const float rmax = __ldg(tc.kvref + ss_n * 2), rmin = __ldg(tc.kvref + ss_n * 2 + 1);
      const float2* qp = reinterpret_cast<const float2*>(cb.q_tok + fi_n * 16);
#pragma unroll 1
      while (pairs_next < kF3EpiPairs) {
        if (bar != nullptr && mbar_test(bar, ph)) break;
        const float2 q2 = __ldg(qp + pairs_next);
        const int off = (half * 2 + (pairs_next >> 2)) * 2048 + (pairs_next & 3) * 4;
        f3_feature_pair(kv2, q2.x, q2.y, att_a, att_b, rmax, rmin, reinterpret_cast<uint32_t*>(row_a1h + off),
                        reinterpret_cast<uint32_t*>(row_a1l + off));
        ++pairs_next;
      }
    };
    int cmin = row_in_tile ? c0 : 128, cmax = row_in_tile ? c0 + S : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
      cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    }
    const int ch_lo = max(cmin >> 5, 2 * half), ch_hi = min((cmax + 31) >> 5, 2 * half + 2);  // this thread's Gram chunks

    attn_fill(tile_begin + wg, nullptr, 0u);   // first tile of the slot: nothing to hide behind yet
    for (int64_t tile = tile_begin + wg; tile < tile_end; tile += 2) {
      attn_fill(tile, nullptr, 0u);            // whatever the MMA waits of the previous tile left over
      pairs_next = 0;
      const int64_t pos = tiles_per_cell > 1 ? tile / tiles_per_cell : tile * cells_per_tile + lc;
      const int sr = tiles_per_cell > 1 ? (int)(tile % tiles_per_cell) * kTcRows + r : s;
      const bool valid = tiles_per_cell > 1 ? (sr < S && pos < n_cells) : (row_in_tile && pos < n_cells);
      const int64_t cell = valid ? (gp.order ? (int64_t)gp.order[pos] : pos) : 0;
      const int code = (valid && gp.n_keys) ? gp.comp_sorted[pos] : -2;
      const int64_t fi = cb.per_row ? (valid ? cell * S + sr : 0) : cell;
      const int64_t kk = (tile - tile_begin) >> 1;
      F3_STAMP(wg, kk, 0);

      // ---- A3 static part: column 32 = 1.0, columns 33.. = u_ln[0..Lq), then zeros (GEMM5 of the previous tile is done)
      if (half == 1) {
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
      fence_proxy_async();
      tc_fence_before();
      ep_arrive();                                            // E0: D1 / D2 of the slot are free, A3 static part written
      F3_STAMP(wg, kk, 1);
      // ================= GEMM1 -> epilogue 1 =================
      mbar_wait(bar_mma, phase); phase ^= 1;
      tc_fence_after();
      F3_STAMP(wg, kk, 2);
#if F3_LN_BARRIER
      ep128(t_lane, half, r, wg, g1, b1, xch);
#else
      ep128r(t_lane, half, g1, b1);
#endif
      F3_STAMP(wg, kk, 3);
      tc_fence_before();
      ep_arrive();                                            // E1
      // ================= GEMM2 -> epilogue 2 =================
      mbar_wait(bar_mma, phase); phase ^= 1;
      tc_fence_after();
      F3_STAMP(wg, kk, 4);
#if F3_LN_BARRIER
      ep32x(t_d2_lane, half, r, wg, g2, b2, xch, row_a3h, row_a3l);
#else
      ep32xr(t_d2_lane, half, g2, b2, row_a3h, row_a3l);
#endif
      F3_STAMP(wg, kk, 5);
      tc_fence_before();
      fence_proxy_async();
      ep_arrive();                                            // E2
      // ================= GEMM3 -> epilogue 3 =================  (GEMM2 is done: A1 of the slot is free for the next tile)
      attn_fill(tile + 2, bar_mma, phase);
      mbar_wait(bar_mma, phase); phase ^= 1;
      tc_fence_after();
      F3_STAMP(wg, kk, 6);
#if F3_LN_BARRIER
      ep128(t_lane, half, r, wg, g3, b3, xch);
#else
      ep128r(t_lane, half, g3, b3);
#endif
      F3_STAMP(wg, kk, 7);
      tc_fence_before();
      ep_arrive();                                            // E3
      // ================= GEMM4 -> epilogue 4: A5 = [t4 | 1 | ...] over the t2 chunks of A3 =================
      attn_fill(tile + 2, bar_mma, phase);
      mbar_wait(bar_mma, phase); phase ^= 1;
      tc_fence_after();
      F3_STAMP(wg, kk, 8);
#if F3_LN_BARRIER
      ep32x(t_d2_lane, half, r, wg, g4, b4, xch, row_a3h, row_a3l);
#else
      ep32xr(t_d2_lane, half, g4, b4, row_a3h, row_a3l);
#endif
      F3_STAMP(wg, kk, 9);
      tc_fence_before();
      fence_proxy_async();
      ep_arrive();                                            // E4
      // ================= GEMM5 -> residual rows =================
      attn_fill(tile + 2, bar_mma, phase);
      mbar_wait(bar_mma, phase); phase ^= 1;
      tc_fence_after();
      F3_STAMP(wg, kk, 10);
      const int ncol = tc.N5 >> 1, col0 = half * ncol;
      float v[32];
      if (ncol == 16) tmem_ld16(t_lane + col0, v); else tmem_ld32(t_lane + col0, v);
      tmem_ld_wait();
      tc_fence_before();
#pragma unroll
      for (int l = 0; l < 32; ++l) v[l] = (l < ncol && col0 + l < L && valid) ? v[l] : 0.f;
      if (cb.per_row && valid) {
        const float* zb = cb.zbase + fi * L;
#pragma unroll
        for (int l = 0; l < 32; ++l)
          if (l < ncol && col0 + l < L) v[l] += __ldg(zb + col0 + l);
      }
      if (z_out != nullptr && valid) {   // z = z_base + residual, straight from the registers (this thread: 16 / 32 consecutive columns)
        const float* zb = cb.zbase + fi * L;
        float* dst = z_out + (cell * S + sr) * (int64_t)L;
        const bool add_zb = !cb.per_row;
        if ((L & 1) == 0) {
#pragma unroll
          for (int l = 0; l < 32; l += 2)
            if (l < ncol && col0 + l < L) {
              float2 o = make_float2(v[l], v[l + 1]);
              if (add_zb) { const float2 b2v = __ldg(reinterpret_cast<const float2*>(zb + col0 + l)); o.x += b2v.x; o.y += b2v.y; }
              *reinterpret_cast<float2*>(dst + col0 + l) = o;
            }
        } else {
#pragma unroll
          for (int l = 0; l < 32; ++l)
            if (l < ncol && col0 + l < L) dst[col0 + l] = v[l] + (add_zb ? __ldg(zb + col0 + l) : 0.f);
        }
      }
      if (!need_dist) continue;   // (the next tile's E0 tells the issuer that D5 has been read)

      // ================= reference row per cell, norms, (hi, lo) Gram operands (Kz = 32) =================
      if (row_in_tile && s == 0) {
#pragma unroll
        for (int l = 0; l < 16; ++l) cref_s[lc * 32 + col0 + l] = v[l];
      }
      wg_sync(wg);
      {
        const float* crow = cref_s + (row_in_tile ? lc : 0) * 32 + col0;
        float nrm_p = 0.f;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int j2 = 0; j2 < 4; ++j2) {
            const int k0 = cc * 8 + 2 * j2;
            const float2 cr = *reinterpret_cast<const float2*>(crow + k0);
            split2(v[k0] - cr.x, v[k0 + 1] - cr.y, hw[j2], lw[j2]);
            // norms from the values the tensor core sees (hi + lo), so that n_i + n_j - 2 g cancels consistently
            const float ar = h2f_lo(hw[j2]) + h2f_lo(lw[j2]), br = h2f_hi(hw[j2]) + h2f_hi(lw[j2]);
            nrm_p = fmaf(ar, ar, nrm_p);
            nrm_p = fmaf(br, br, nrm_p);
          }
          *reinterpret_cast<uint4*>(zh + (half * 2 + cc) * 2048 + r * 16) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(zl + (half * 2 + cc) * 2048 + r * 16) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        xch[half * 128 + r].x = nrm_p;
      }
      fence_proxy_async();
      ep_arrive();                                            // E5: Gram operands in place
      F3_STAMP(wg, kk, 11);
      wg_sync(wg);                                            // norm halves visible
      const float n_i = xch[r].x + xch[128 + r].x;
      if (half == 0) nrm_s[r] = n_i;
      attn_fill(tile + 2, bar_mma, phase);
      mbar_wait(bar_mma, phase); phase ^= 1;
      tc_fence_after();

      F3_STAMP(wg, kk, 12);
      // ================= group-run bookkeeping =================
      bool use_acc = false;
      if (gp.n_keys > 0) {
        const int first_code = gp.comp_sorted[min(tile * cells_per_tile, n_cells - 1)];
        use_acc = wg_all(wg, ((!valid) || code == first_code) ? 1 : 0) != 0;  // also publishes nrm_s
        if (use_acc) {
          if (leader) { while (atomicCAS(&ctrl->lock, 0, 1) != 0) __nanosleep(20); }
          wg_sync(wg);
          tc_fence_after();
          const int cur = *reinterpret_cast<volatile int*>(&ctrl->cur_code);
          const int cur_valid = *reinterpret_cast<volatile int*>(&ctrl->cur_valid);
          if (cur_valid && cur != first_code) {
            const long long rep = *reinterpret_cast<volatile long long*>(&ctrl->cur_cell);
            flush_acc(t_acc_lane, half, gp, rep, S, c0, row_in_tile, r);
          }
          wg_sync(wg);
          if (leader) { ctrl->cur_code = first_code; ctrl->cur_valid = 1; ctrl->cur_cell = gp.order[tile * cells_per_tile]; }
        }
      } else {
        wg_sync(wg);  // publishes nrm_s
      }
      F3_STAMP(wg, kk, 13);
      // ================= distance epilogue: this thread's chunks of row r =================
#pragma unroll 1
      for (int c = ch_lo; c < ch_hi; ++c) {
        float g[32];
        tmem_ld32(t_lane + c * 32, g);
        tmem_ld_wait();
        uint32_t flagged = 0u;
        const bool full = __all_sync(0xffffffffu, valid && c * 32 >= c0 && c * 32 + 32 <= c0 + S);
        const float2 ni2 = make_float2(n_i, n_i), m2 = make_float2(-2.0f, -2.0f), thr2 = make_float2(-1.0f / 4096.0f, -1.0f / 4096.0f);
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 nj = *reinterpret_cast<const float4*>(nrm_s + c * 32 + 4 * q4);
          const float2 njp[2] = {make_float2(nj.x, nj.y), make_float2(nj.z, nj.w)};
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int q = 4 * q4 + 2 * e2, j = c * 32 + q;
            const float2 sn = __fadd2_rn(ni2, njp[e2]);
            const float2 d2 = __ffma2_rn(make_float2(g[q], g[q + 1]), m2, sn);
            const float2 tt = __ffma2_rn(sn, thr2, d2);        // < 0: cancellation dominated
            const bool in0 = full ? (j != r) : (valid && (unsigned)(j - c0) < (unsigned)S && j != r);
            const bool in1 = full ? (j + 1 != r) : (valid && (unsigned)(j + 1 - c0) < (unsigned)S && j + 1 != r);
            if (in0 && tt.x < 0.f) flagged |= 1u << q;
            if (in1 && tt.y < 0.f) flagged |= 2u << q;
            g[q] = in0 ? sqrt_approx(fmaxf(d2.x, 0.f)) : 0.f;
            g[q + 1] = in1 ? sqrt_approx(fmaxf(d2.y, 0.f)) : 0.f;
          }
        }
        if (__any_sync(0xffffffffu, flagged != 0u))  // rare: near-duplicate samples -> exact direct-difference distance
          exact_fix_warp<2048u>(g, flagged, zh, zl, r, c * 32, L);
        if (cell_out != nullptr && (S & 3) == 0) {
          // sector-coalesced stores through a per-warp shared-memory tile (8 columns = 32 bytes per row and sub-step;
          // 16-byte units XOR-swizzled with the row: conflict-free on both sides); every store instruction writes full
          // 32-byte sectors of 16 rows.  Rows of a warp may belong to different cells: destination row and first
          // column of the cell travel by shuffle.
          float4* T = reinterpret_cast<float4*>(slot + sm.off_T) + (wt >> 5) * 64;
          const int my_row = valid ? (int)(cell * S + s) : -1;   // < 2^31: cells are indices inside one device chunk
#pragma unroll
          for (int h4 = 0; h4 < 4; ++h4) {
            const int sw = (lane >> 2) & 1;
            T[lane * 2 + (0 ^ sw)] = make_float4(g[h4 * 8], g[h4 * 8 + 1], g[h4 * 8 + 2], g[h4 * 8 + 3]);
            T[lane * 2 + (1 ^ sw)] = make_float4(g[h4 * 8 + 4], g[h4 * 8 + 5], g[h4 * 8 + 6], g[h4 * 8 + 7]);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 2; ++it) {
              const int rl = it * 16 + (lane >> 1), unit = lane & 1;
              const int orow = __shfl_sync(0xffffffffu, my_row, rl);
              const int oc0 = __shfl_sync(0xffffffffu, c0, rl);
              const int j = c * 32 + h4 * 8 + unit * 4 - oc0;
              const float4 val = T[rl * 2 + (unit ^ ((rl >> 2) & 1))];
              if (orow >= 0 && j >= 0 && j < S) *reinterpret_cast<float4*>(cell_out + (size_t)orow * S + j) = val;
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
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {   // 16 columns at a time: 16 live registers instead of 32 next to g[32]
            float acc[16];
            tmem_ld16(t_acc_lane + c * 32 + hh * 16, acc);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float2 t = __fadd2_rn(make_float2(acc[2 * q], acc[2 * q + 1]), make_float2(g[hh * 16 + 2 * q], g[hh * 16 + 2 * q + 1]));
              acc[2 * q] = t.x; acc[2 * q + 1] = t.y;
            }
            tmem_st16(t_acc_lane + c * 32 + hh * 16, reinterpret_cast<uint32_t*>(acc));
          }
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
      F3_STAMP(wg, kk, 14);
      if (use_acc) {
        tmem_st_wait();
        tc_fence_before();
        wg_sync(wg);
        if (leader) { __threadfence_block(); atomicExch(&ctrl->lock, 0); }
      }
      tc_fence_before();
      wg_sync(wg);  // Gram operands / staging tiles / nrm_s fully consumed before the next tile's A3 writes
      F3_STAMP(wg, kk, 15);
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (gp.n_keys > 0 && warp < 8 && ctrl->cur_valid) {
    const int wt = tid & 255, half = wt >> 7, r = wt & 127;
    const int lc = r / S, c0 = lc * S;
    const uint32_t t_acc_lane = tmem_base + 384 + ((uint32_t)(((wt >> 5) & 3) * 32) << 16);
    flush_acc(t_acc_lane, half, gp, ctrl->cur_cell, S, c0, lc < cells_per_tile, r);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// fused distances: 4 <= S <= 128, l2, n_latent <= 32; chain-only (z output): any S / n_latent the chain supports
inline bool tc_fused3_supported(const TcNet& tc, int S, int L, int norm) {
  return tc.ready && tc.fused3_smem_bytes != 0 && norm == MRVI_NORM_L2 && S >= 4 && S <= 128 && L <= 32;
}

inline int tc_fused3_prepare(TcNet& tc) {
  const size_t bytes = f3_smem_map(tc.wimg_bytes, tc.K1, tc.K3).total + 1024;
  if (bytes > (size_t)kMaxSmemOptin) return 0;  // unavailable for this shape: the second version / unfused kernels are used
  if (cudaFuncSetAttribute(k_chain_fused3_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin) != cudaSuccess) {
    g_tc_error = "cudaFuncSetAttribute(k_chain_fused3_tc) failed";
    return MRVI_ERR_CUDA;
  }
  tc.fused3_smem_bytes = bytes;
  return 0;
}

inline void tc_launch_fused3(const TcNet& tc, const NetW& net, int64_t n, const CellBuf& cb, float* z_out, float* cell_out,
                             const GroupPlan& gp, cudaStream_t st) {
  const int S = net.S;
  int cpt = 1, tpc = 1;
  int64_t n_tiles;
  if (S <= kTcRows) { cpt = kTcRows / S; n_tiles = (n + cpt - 1) / cpt; }
  else { tpc = (S + kTcRows - 1) / kTcRows; n_tiles = n * tpc; }   // chain only
  if (n_tiles == 0) return;
  const int grid = (int)std::min<int64_t>((n_tiles + 1) / 2, tc.sm_count);
  const int64_t tiles_per_cta = (n_tiles + grid - 1) / grid;
  k_chain_fused3_tc<<<grid, kF3Threads, tc.fused3_smem_bytes, st>>>(tc, cb, n, S, net.L, net.Lq, z_out, cell_out, gp, n_tiles, cpt,
                                                                    tiles_per_cta, tpc);
}

}  // namespace mrvi
