// Fused tensor-core kernel, fourth version: q(z|u,s) chain + pairwise distances + group reduction, PING-PONG over two
// tile slots (round 2).  Same arithmetic and operand formats as k_chain_fused_tc (tc_fused.cuh); what changed is the
// schedule, driven by the ncu source page of the second version (profiles/r2_fused_c4_ncu.json, tools/ncu_segments.py):
// its two warpgroups each own ONE tile, so every one of the six GEMMs of a tile (issue ~45-62 clk per tcgen05.mma +
// ~500 clk commit -> mbarrier latency, tools/mma_microbench.cu) is waited for by the eight warps that need its
// result -- 17 % of all warp time sat on the `@!P BRA` behind the six mbarrier try_waits.
//
//   * ALL sixteen epilogue warps work on one tile at a time, FOUR threads per row (thread (r, cq) owns column quarter cq
//     of every epilogue), and alternate between the two slots phase by phase: while they run epilogue k of slot A the
//     tensor core runs GEMM k of slot B.  A GEMM's latency is hidden whenever the other slot's epilogue takes longer.
//   * Warp 16 is the MMA issuer for both slots (warp-uniform branch + elect.sync, descriptors in uniform registers); it
//     walks the fixed sequence G1(A) G1(B) G2(A) G2(B) ... behind the slots' "operands ready" mbarriers (16 warp
//     arrivals each) and commits to the slots' "accumulator ready" mbarriers.  The skip-connection parts of GEMM2 / GEMM4
//     (A1.B2a, A3.B4a) are issued right behind GEMM1 / GEMM3.
//   * No 256 / 512-thread barrier inside the chain: the LayerNorm statistics of a row are exchanged between its four
//     threads behind a 128-thread named barrier (the four warps that share a TMEM lane quarter), double buffered.  The
//     128-wide epilogues read their accumulator columns ONCE (32 values stay in registers across that barrier).
//   * The group accumulator needs no lock: one tile is in the distance epilogue at a time, every thread owns its
//     (lane, column quarter) of the 128 x 128 TMEM accumulator, and every thread tracks the run state itself (tiles are
//     visited in sorted order, so the sequence of composite codes is the same for all of them).
//   * Registers follow the roles (`setmaxnreg`): 104 for the epilogue warps, 64 for the issuer's warpgroup.
//
// Block = 640 threads: warps 0-15 epilogue (warp = 4 cq + q: lane quarter q = rows [32 q, 32 q + 32), column quarter cq),
// warp 16 issuer, warps 17-19 idle (warps are allocated in fours).  TMEM: slot s: [192 s, +128) D1/T/D3/T/D5/Gram,
// [192 s + 128, +64) D2/D4 as [A.W_hi | A.W_lo]; [384, 512) group accumulator.
#pragma once
#include "tc_fused3.cuh"

namespace mrvi {

#ifdef MRVI_DBG_F4
// clock64 timeline of CTA 0 (tools/dbg/run_dbg_f4.py): epilogue warps 0 and 15: [which][pair < 16][phase < 7][slot][before wait, start, end, -];
// issuer: [pair < 16][step < 6][slot][ready seen, issued]
__device__ long long g_dbg_f4[2 * 16 * 7 * 2 * 4 + 16 * 6 * 2 * 2];
#define F4_STAMP(k_) do { if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 15) && pair_i < 16) g_dbg_f4[(((((warp == 15) * 16 + pair_i) * 7 + ph) * 2 + sl) * 4) + (k_)] = clock64(); } while (0)
#define F4_ISTAMP(k_) do { if (blockIdx.x == 0 && lane == 0 && pair_i < 16) g_dbg_f4[1792 + (((pair_i * 6 + step) * 2 + s) * 2) + (k_)] = clock64(); } while (0)
#else
#define F4_STAMP(k_) do { } while (0)
#define F4_ISTAMP(k_) do { } while (0)
#endif

constexpr int kF4Threads = 640;
constexpr int kF4RegsEpi = 104, kF4RegsAux = 64;

struct F4Smem {
  uint32_t w;                               // weight image
  uint32_t slot[2];                         // [a1h | a1l (4 feature chunks) | a3h | a3l], overlaid after GEMM5 by [zf] / [zh zl | cref | T]
  uint32_t off_a1l, off_a3h, off_a3l;       // relative to the slot base (a1h at 0)
  uint32_t off_zf, off_zh, off_zl, off_cref, off_T;
  uint32_t nrm[2];                          // [128] fp32 row norms of the slot's Gram operands
  uint32_t xch;                             // [2 buffers][4 quarters][128 rows] float2: partial row statistics
  uint32_t lnv, ctrl, total;
  int zf_ld, Kz;
};

__host__ __device__ inline F4Smem f4_smem_map(uint32_t wimg_bytes, int K1, int K3, int L) {
  F4Smem m;
  uint32_t o = 0;
  auto up = [](uint32_t b) { return (b + 1023u) & ~1023u; };
  m.w = o; o += up(wimg_bytes);
  const uint32_t a1 = up(128 * K1 * 2), a1l = 128 * 32 * 2, a3 = up(128 * K3 * 2);
  m.off_a1l = a1; m.off_a3h = a1 + a1l; m.off_a3l = a1 + a1l + a3;
  const uint32_t abytes = a1 + a1l + 2 * a3;
  m.Kz = L <= 32 ? 32 : 64;
  m.zf_ld = m.Kz + 1;
  const uint32_t zf = up(128 * m.zf_ld * 4), zk = 128 * m.Kz * 2;
  m.off_zf = 0; m.off_zh = 0; m.off_zl = zk; m.off_cref = 2 * zk; m.off_T = 2 * zk + 32u * m.Kz * 4u;
  uint32_t scratch = abytes;
  if (scratch < zf + 2048u) scratch = zf + 2048u;           // fp32 z tile + row info (z output only)
  if (scratch < m.off_T + 16384u) scratch = m.off_T + 16384u;  // keep_cell staging: 16 warps x 1 KB
  m.slot[0] = o; o += scratch;
  m.slot[1] = o; o += scratch;
  m.nrm[0] = o; o += 512;
  m.nrm[1] = o; o += 512;
  m.xch = o; o += 2 * 4 * 128 * 8;
  m.lnv = o; o += (uint32_t)((sizeof(TcLn) + 127) & ~size_t(127));
  m.ctrl = o; o += 128;
  m.total = o;
  return m;
}

struct F4Ctrl {
  uint64_t bar_w;
  uint64_t ready[2];   // operands of the slot's next GEMM are in place (16 warp arrivals)
  uint64_t done[2];    // the slot's GEMM has completed (tcgen05.commit)
  uint32_t tmem_slot;
};

__device__ __forceinline__ void q_sync(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory"); }   // the 4 warps of a lane quarter
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 5, 512;" ::: "memory"); }                 // the 16 epilogue warps

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

// Row statistics of a row from the partial sums of its four threads (fixed order: the four threads get bit-identical
// values).  `xb`: buffer of this exchange; the caller alternates it.
__device__ __forceinline__ float2 f4_exchange(float2* xch, int xb, int cq, int r, int q, float2 mine) {
  float2* buf = xch + xb * 512;
  buf[cq * 128 + r] = mine;
  q_sync(q);
  const float2 p0 = buf[r], p1 = buf[128 + r], p2 = buf[256 + r], p3 = buf[384 + r];
  return make_float2((p0.x + p1.x) + (p2.x + p3.x), (p0.y + p1.y) + (p2.y + p3.y));
}

// LayerNorm + gelu of the 128-column accumulator row, this thread's 32-column quarter, written back in place as
// [16 hi words | 16 lo words]: ONE TMEM read, the values stay in registers across the statistics exchange.
__device__ __forceinline__ void ep128q(uint32_t t_lane, int cq, int r, int q, const float* g_s, const float* b_s, float2* xch, int xb) {
  float v[32];
  tmem_ld32(t_lane + cq * 32, v);
  tmem_ld_wait();
  float2 a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 x = make_float2(v[2 * i], v[2 * i + 1]);
    a1 = __fadd2_rn(a1, x);
    a2 = __ffma2_rn(x, x, a2);
  }
  const float2 st = f4_exchange(xch, xb, cq, r, q, make_float2(a1.x + a1.y, a2.x + a2.y));
  const float mean = st.x * (1.0f / 128);
  const float var = fmaxf(fmaf(st.y, 1.0f / 128, -mean * mean), 0.f);
  const float rstd = rsqrt_approx(var + kLnEps);
  const float2 rs2 = make_float2(rstd, rstd), nb2 = make_float2(-mean * rstd, -mean * rstd);
  uint32_t out[32];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 g4 = *reinterpret_cast<const float4*>(g_s + cq * 32 + 4 * k);
    const float4 b4 = *reinterpret_cast<const float4*>(b_s + cq * 32 + 4 * k);
    float2 x0 = __ffma2_rn(make_float2(v[4 * k], v[4 * k + 1]), rs2, nb2);
    float2 x1 = __ffma2_rn(make_float2(v[4 * k + 2], v[4 * k + 3]), rs2, nb2);
    x0 = gelu_fast2(__ffma2_rn(x0, make_float2(g4.x, g4.y), make_float2(b4.x, b4.y)));
    x1 = gelu_fast2(__ffma2_rn(x1, make_float2(g4.z, g4.w), make_float2(b4.z, b4.w)));
    split2v(x0, out[2 * k], out[16 + 2 * k]);
    split2v(x1, out[2 * k + 1], out[16 + 2 * k + 1]);
  }
  tmem_st32(t_lane + cq * 32, out);
  tmem_st_wait();
}

// LayerNorm + gelu of the 32-wide accumulator stored as two halves [A.W_hi | A.W_lo] (columns [0,32) / [32,64) of the D2
// region): this thread's 8 columns -> one 16-byte K chunk (hi and lo copy) of the next A operand in shared memory.
__device__ __forceinline__ void ep32xq(uint32_t t_d2_lane, int cq, int r, int q, const float* g_s, const float* b_s, float2* xch, int xb,
                                       uint8_t* row_h, uint8_t* row_l) {
  float v[8], w[8];
  tmem_ld8(t_d2_lane + cq * 8, v);
  tmem_ld8(t_d2_lane + 32 + cq * 8, w);
  tmem_ld_wait();
  float2 a1 = make_float2(0.f, 0.f), a2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 x = __fadd2_rn(make_float2(v[2 * i], v[2 * i + 1]), make_float2(w[2 * i], w[2 * i + 1]));
    v[2 * i] = x.x; v[2 * i + 1] = x.y;
    a1 = __fadd2_rn(a1, x);
    a2 = __ffma2_rn(x, x, a2);
  }
  const float2 st = f4_exchange(xch, xb, cq, r, q, make_float2(a1.x + a1.y, a2.x + a2.y));
  const float mean = st.x * (1.0f / 32);
  const float var = fmaxf(fmaf(st.y, 1.0f / 32, -mean * mean), 0.f);
  const float rstd = rsqrt_approx(var + kLnEps);
  const float2 rs2 = make_float2(rstd, rstd), nb2 = make_float2(-mean * rstd, -mean * rstd);
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 gg = *reinterpret_cast<const float2*>(g_s + cq * 8 + 2 * j);
    const float2 bb = *reinterpret_cast<const float2*>(b_s + cq * 8 + 2 * j);
    float2 x = __ffma2_rn(make_float2(v[2 * j], v[2 * j + 1]), rs2, nb2);
    x = gelu_fast2(__ffma2_rn(x, gg, bb));
    split2v(x, h[j], l[j]);
  }
  *reinterpret_cast<uint4*>(row_h + cq * 2048) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(row_l + cq * 2048) = make_uint4(l[0], l[1], l[2], l[3]);
}

// flush this thread's 32-column quarter of its accumulator lane into the float64 group sums and clear it
__device__ __forceinline__ void f4_flush_acc(uint32_t t_acc_lane, int cq, const GroupPlan& gp, long long rep_cell, int S, int c0,
                                             bool lane_valid, int r) {
  float a[32];
  tmem_ld32(t_acc_lane + cq * 32, a);
  tmem_ld_wait();
  if (lane_valid) {
#pragma unroll 1
    for (int k = 0; k < gp.n_keys; ++k) {
      const int g = gp.codes[k][rep_cell];
      if (g < 0 || g >= gp.n_groups[k]) continue;
      double* base = gp.sums[k] + ((size_t)g * S + (r - c0)) * S;
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const int j = cq * 32 + e - c0;
        if (j >= 0 && j < S && j != r - c0) atomicAdd(base + j, (double)a[e]);
      }
    }
  }
  uint32_t z[32];
#pragma unroll
  for (int e = 0; e < 32; ++e) z[e] = 0u;
  tmem_st32(t_acc_lane + cq * 32, z);
  tmem_st_wait();
}

__global__ void __launch_bounds__(kF4Threads, 1)
k_chain_fused4_tc(const __grid_constant__ TcNet tc, CellBuf cb, int64_t n_cells, int S, int L, int Lq,
                  float* __restrict__ z_out, float* __restrict__ cell_out, GroupPlan gp, int64_t n_tiles,
                  int cells_per_tile, int64_t tiles_per_cta, int tiles_per_cell) {
  extern __shared__ __align__(1024) uint8_t smem_tc[];
  uint8_t* const smem = smem_tc;
  const F4Smem sm = f4_smem_map(tc.wimg_bytes, tc.K1, tc.K3, L);
  F4Ctrl* ctrl = reinterpret_cast<F4Ctrl*>(smem + sm.ctrl);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
  const int lane = tid & 31;
  float* lnv = reinterpret_cast<float*>(smem + sm.lnv);
  const bool need_dist = (cell_out != nullptr) || (gp.n_keys > 0);
  const int nsteps = need_dist ? 6 : 5;   // GEMMs per tile

  if (tid == 0) {
    mbar_init(&ctrl->bar_w, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&ctrl->ready[s], 16); mbar_init(&ctrl->done[s], 1); }
    fence_mbar_init();
  }
  for (int i = tid; i < (int)(sizeof(TcLn) / sizeof(float)); i += kF4Threads) lnv[i] = reinterpret_cast<const float*>(&tc.ln)[i];
  if (warp == 0) tmem_alloc(&ctrl->tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctrl->tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(&ctrl->bar_w, tc.wimg_bytes);
    bulk_g2s(smem + sm.w, tc.wimg, tc.wimg_bytes, &ctrl->bar_w);
  }
  if (gp.n_keys > 0 && warp < 16) {  // clear the group accumulator once: every thread its own (lane, column quarter)
    uint32_t z[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) z[e] = 0u;
    tmem_st32(tmem_base + 384 + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 32, z);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int64_t tile_begin = (int64_t)blockIdx.x * tiles_per_cta;
  const int64_t tile_end = min(n_tiles, tile_begin + tiles_per_cta);

  if (warp >= 16) {
    setmaxnreg_dec<kF4RegsAux>();
    if (warp == 16) {
      // ============================================ MMA issuer ============================================
      const uint32_t s_w = smem_u32(smem + sm.w);
      const uint32_t idesc128 = umma_idesc_f16(128), idesc64 = umma_idesc_f16(64), idesc32 = umma_idesc_f16(32),
                     idescN5 = umma_idesc_f16(tc.N5);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t phr = 0;   // parity bits of ready[0], ready[1]
      mbar_wait(&ctrl->bar_w, 0);
      for (int64_t t0 = tile_begin; t0 < tile_end; t0 += 2) {
        const bool has_b = t0 + 1 < tile_end;
        const int64_t pair_i = (t0 - tile_begin) >> 1; (void)pair_i;
#pragma unroll 1
        for (int step = 0; step < nsteps; ++step) {
#pragma unroll 1
          for (int s = 0; s < 2; ++s) {
            if (s && !has_b) break;
            mbar_wait(&ctrl->ready[s], (phr >> s) & 1u);
            phr ^= 1u << s;
            tc_fence_after();
            F4_ISTAMP(0);
            const uint32_t t_d1 = tb + s * 192, t_d2 = t_d1 + 128;
            const uint32_t sa1h = smem_u32(smem) + sm.slot[0] + (uint32_t)s * (sm.slot[1] - sm.slot[0]), sa1l = sa1h + sm.off_a1l, sa3h = sa1h + sm.off_a3h, sa3l = sa1h + sm.off_a3l;
            if (elect_one()) {
              if (step == 0) {
                f3_gemm_ss128(t_d1, sa1h, sa1l, s_w, tc.b1, tc.K1, 32, idesc128);
                umma_commit(&ctrl->done[s]);
                f3_gemm_ss_cat(t_d2, sa1h, sa1l, s_w, tc.b2a, tc.K1, 32, idesc64, idesc32);   // skip part of GEMM2, under epilogue 1
              } else if (step == 1) {
                f3_gemm_ts_cat(t_d2, t_d1, s_w, tc.b2t, idesc64, idesc32);
                umma_commit(&ctrl->done[s]);
              } else if (step == 2) {
                f3_gemm_ss128(t_d1, sa3h, sa3l, s_w, tc.b3, tc.K3, tc.K3, idesc128);
                umma_commit(&ctrl->done[s]);
                f3_gemm_ss_cat(t_d2, sa3h, sa3l, s_w, tc.b4a, tc.K3, tc.K3, idesc64, idesc32);  // skip part of GEMM4, under epilogue 3
              } else if (step == 3) {
                f3_gemm_ts_cat(t_d2, t_d1, s_w, tc.b4t, idesc64, idesc32);
                umma_commit(&ctrl->done[s]);
              } else if (step == 4) {
                f3_gemm_ss128(t_d1, sa1h, sa1l, s_w, tc.b5, tc.K5, 32, idescN5);                // A5 = [t4 | 1 | 0] in the A1 buffers
                umma_commit(&ctrl->done[s]);
              } else {
                // Gram: the two cross terms first, Zh.Zh^T last (see tc_fused.cuh: near-symmetric output)
                const uint32_t zh = sa1h + sm.off_zh, zl = sa1h + sm.off_zl;
                uint64_t dh = umma_desc(zh, 2048, 128), dl = umma_desc(zl, 2048, 128);
                const uint64_t dh0 = dh;
                for (int ks = 0; ks < sm.Kz / 16; ++ks) {
                  umma_f16(t_d1, dh, dl, idesc128, ks == 0 ? 0u : 1u);
                  umma_f16(t_d1, dl, dh, idesc128, 1u);
                  dh = desc_adv(dh, 4096); dl = desc_adv(dl, 4096);
                }
                dh = dh0;
                for (int ks = 0; ks < sm.Kz / 16; ++ks) {
                  umma_f16(t_d1, dh, dh, idesc128, 1u);
                  dh = desc_adv(dh, 4096);
                }
                umma_commit(&ctrl->done[s]);
              }
            }
            __syncwarp();
            F4_ISTAMP(1);
          }
        }
      }
    }
  } else {
    // ============================================ epilogue warps ============================================
    setmaxnreg_inc<kF4RegsEpi>();
    const int q = warp & 3, cq = warp >> 2, r = q * 32 + lane;
    const float *g1 = lnv, *b1 = lnv + 128, *g2 = lnv + 256, *b2 = lnv + 288, *g3 = lnv + 320, *b3 = lnv + 448,
                *g4 = lnv + 576, *b4 = lnv + 608;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t t_acc_lane = tmem_base + 384 + lane_off;
    float2* xch = reinterpret_cast<float2*>(smem + sm.xch);
    const int Kz = sm.Kz, zld = sm.zf_ld;
    int xb = 0;          // statistics-exchange buffer (alternates: a thread is never more than one exchange ahead of its row's other threads)
    uint32_t phd = 0;    // parity bits of done[0], done[1]
    // static row -> (local cell, sample) map of this thread
    const int lc = r / S, s = r - lc * S, c0 = lc * S;
    const bool row_in_tile = lc < cells_per_tile;
    // does this warp's column quarter hold Gram entries of its rows' cells?  (rows of a warp may straddle cells)
    int cmin = row_in_tile ? c0 : 128, cmax = row_in_tile ? c0 + S : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
      cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    }
    const bool chunk_needed = cq >= (cmin >> 5) && cq < ((cmax + 31) >> 5);
    const int head = cq >> 1;
    const float a_h = head ? tc.alpha[1] : tc.alpha[0], b_h = head ? tc.beta[1] : tc.beta[0];
    // group-run state, tracked identically by every thread
    int cur_code = -1;
    bool cur_valid = false;
    long long cur_cell = 0;

    mbar_wait(&ctrl->bar_w, 0);   // (the weight image is only read by the MMAs, but the wait keeps the first arrivals behind the copy)
    const int nph = nsteps + 1;
    for (int64_t t0 = tile_begin; t0 < tile_end; t0 += 2) {
      const bool has_b = t0 + 1 < tile_end;
      const int64_t pair_i = (t0 - tile_begin) >> 1; (void)pair_i;
      epi_sync();   // the previous pair is fully consumed (Gram operands, reference rows, norms of both slots)
#pragma unroll 1
      for (int ph = 0; ph < nph; ++ph) {
#pragma unroll 1
        for (int sl = 0; sl < 2; ++sl) {
          if (sl && !has_b) break;
          const int64_t tile = t0 + sl;
          uint8_t* slot = smem + sm.slot[0] + (uint32_t)sl * (sm.slot[1] - sm.slot[0]);
          const uint32_t t_lane = tmem_base + sl * 192 + lane_off, t_d2_lane = t_lane + 128;
          F4_STAMP(0);
          if (ph > 0) {
            mbar_wait(&ctrl->done[sl], (phd >> sl) & 1u);
            phd ^= 1u << sl;
            tc_fence_after();
          }
          F4_STAMP(1);
          // row -> cell of this tile (tiles_per_cell > 1: S > 128, chain only: tile -> (cell, 128-row slice of its samples))
          const int64_t pos = tiles_per_cell > 1 ? tile / tiles_per_cell : tile * cells_per_tile + lc;
          const int sr = tiles_per_cell > 1 ? (int)(tile % tiles_per_cell) * kTcRows + r : s;
          const bool valid = tiles_per_cell > 1 ? (sr < S && pos < n_cells) : (row_in_tile && pos < n_cells);

          if (ph == 0) {
            // ================= attention features + static operand columns =================
            const int64_t cell = valid ? (gp.order ? (int64_t)gp.order[pos] : pos) : 0;
            const int ss = valid ? sr : 0;
            const int64_t fi = cb.per_row ? (valid ? cell * S + sr : 0) : cell;
            uint8_t* row_a1h = slot + r * 16;
            uint8_t* row_a1l = slot + sm.off_a1l + r * 16;
            uint32_t hw[4], lw[4];
            if (tc.att_tab != nullptr && !cb.per_row) {
              // table path (tc_build): 8 look-ups of this row's sample table; the rows of a cell share the cell's query tokens,
              // so the lanes of a warp read consecutive 16-byte entries of one interval
              const float4* tbp = tc.att_tab + ss;
              const float4* qp4 = reinterpret_cast<const float4*>(cb.q_tok + fi * 16) + (cq & 1) * 2;
              const float t0s = -tc.att_t0 * tc.att_scale;
              const float as = a_h * tc.att_scale, bs = fmaf(b_h, tc.att_scale, t0s);   // x = (alpha q + beta - t0) * scale
              const float4 qa = __ldg(qp4), qb = __ldg(qp4 + 1);
              const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
              float4 nd[8];
              float fr[8];
#pragma unroll
              for (int f = 0; f < 8; ++f) {
                const float x = fminf(fmaxf(fmaf(as, qv[f], bs), 0.f), tc.att_xmax);
                const int k = (int)x;
                fr[f] = x - (float)k;
                nd[f] = __ldg(tbp + (size_t)k * S);
              }
#pragma unroll
              for (int f2 = 0; f2 < 4; ++f2) {
                float mm[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const float4 n4 = nd[2 * f2 + e];
                  const float u = fr[2 * f2 + e], df = n4.z - n4.x;
                  const float c3 = fmaf(-2.f, df, n4.y + n4.w), c2 = fmaf(3.f, df, fmaf(-2.f, n4.y, -n4.w));
                  mm[e] = fmaf(u, fmaf(u, fmaf(u, c3, c2), n4.y), n4.x);
                }
                split2(mm[0], mm[1], hw[f2], lw[f2]);
              }
            } else {
              float2 kv2[8];  // kv * log2(e), pairs
              const float4* kp = reinterpret_cast<const float4*>(tc.kvl + (size_t)ss * 16);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float4 t4 = __ldg(kp + k);
                kv2[2 * k] = make_float2(t4.x, t4.y);
                kv2[2 * k + 1] = make_float2(t4.z, t4.w);
              }
              const float rmax = __ldg(tc.kvref + ss * 2), rmin = __ldg(tc.kvref + ss * 2 + 1);
              const float2* qp = reinterpret_cast<const float2*>(cb.q_tok + fi * 16) + (cq & 1) * 4;
#pragma unroll 1
              for (int it = 0; it < 4; ++it) {  // one packed word (two features) per iteration
                const float2 qc = __ldg(qp + it);
                float mm[2];
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
                uint32_t h1, l1;
                split2(mm[0], mm[1], h1, l1);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (k == it) { hw[k] = h1; lw[k] = l1; }
              }
            }
            *reinterpret_cast<uint4*>(row_a1h + cq * 2048) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(row_a1l + cq * 2048) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            if (cq == 0) {  // constant-one column (k = 32) and zero padding of A1 (the lo copy holds the feature chunks only)
              for (int c = 4; c < tc.K1 / 8; ++c)
                *reinterpret_cast<uint4*>(row_a1h + c * 2048) = make_uint4(c == 4 ? 0x00003C00u : 0u, 0u, 0u, 0u);
            } else {        // A3 static part: column 32 = 1.0, columns 33.. = u_ln[0..Lq), then zeros; chunks dealt to cq = 1, 2, 3
              const float* ul = cb.u_ln + fi * Lq;
              uint8_t* row_a3h = slot + sm.off_a3h + r * 16;
              uint8_t* row_a3l = slot + sm.off_a3l + r * 16;
              for (int c = 4 + (cq - 1); c < tc.K3 / 8; c += 3) {
                uint32_t h3[4], l3[4];
#pragma unroll
                for (int j2 = 0; j2 < 4; ++j2) {
                  const int k0 = (c - 4) * 8 + 2 * j2 - 1, k1 = k0 + 1;   // index into u_ln; -1 = the constant column
                  const float a = k0 < 0 ? 1.0f : (k0 < Lq ? __ldg(ul + k0) : 0.f);
                  const float b = k1 < Lq ? __ldg(ul + k1) : 0.f;
                  split2(a, b, h3[j2], l3[j2]);
                }
                *reinterpret_cast<uint4*>(row_a3h + c * 2048) = make_uint4(h3[0], h3[1], h3[2], h3[3]);
                *reinterpret_cast<uint4*>(row_a3l + c * 2048) = make_uint4(l3[0], l3[1], l3[2], l3[3]);
              }
            }
            fence_proxy_async();
          } else if (ph == 1 || ph == 3) {
            ep128q(t_lane, cq, r, q, ph == 1 ? g1 : g3, ph == 1 ? b1 : b3, xch, xb);
            xb ^= 1;
            tc_fence_before();
          } else if (ph == 2 || ph == 4) {
            // epilogue 2 -> A3 feature chunks; epilogue 4 -> A5 = [t4 | 1 | 0] in the A1 buffers (GEMM2 was their last reader)
            uint8_t* row_h = slot + (ph == 2 ? sm.off_a3h : 0u) + r * 16;
            uint8_t* row_l = slot + (ph == 2 ? sm.off_a3l : sm.off_a1l) + r * 16;
            ep32xq(t_d2_lane, cq, r, q, ph == 2 ? g2 : g4, ph == 2 ? b2 : b4, xch, xb, row_h, row_l);
            xb ^= 1;
            tc_fence_before();
            fence_proxy_async();
          } else if (ph == 5) {
            // ================= residual rows: z output, reference row per cell, norms, (hi, lo) Gram operands =================
            const int64_t cell = valid ? (gp.order ? (int64_t)gp.order[pos] : pos) : 0;
            const int64_t fi = cb.per_row ? (valid ? cell * S + sr : 0) : cell;
            const int ncol = Kz >> 2, col0 = cq * ncol;   // this thread's quarter of the N5 = Kz output columns (8 or 16)
            float v[16];
            if (ncol == 8) tmem_ld8(t_lane + col0, v); else tmem_ld16(t_lane + col0, v);
            tmem_ld_wait();
            // (columns >= L of D5 are exactly zero: the padded rows of the B5 operand are; only rows outside the chunk need zeroing)
#pragma unroll
            for (int l = 0; l < 16; ++l)
              if (!valid || l >= ncol) v[l] = 0.f;
            if (cb.per_row && valid) {
              const float* zb = cb.zbase + fi * L;
#pragma unroll
              for (int l = 0; l < 16; ++l)
                if (l < ncol && col0 + l < L) v[l] += __ldg(zb + col0 + l);
            }
            tc_fence_before();
            if (z_out != nullptr) {
              // A1/A3 are dead (GEMM5 done): stage the fp32 residual tile on them, then write z = z_base + residual with
              // fully coalesced 4-byte stores: consecutive threads walk the tile row-major
              float* zf = reinterpret_cast<float*>(slot + sm.off_zf);
              long long* rowinfo = reinterpret_cast<long long*>(slot + ((128u * (uint32_t)zld * 4u + 1023u) & ~1023u));
#pragma unroll
              for (int l = 0; l < 16; ++l)
                if (l < ncol) zf[r * zld + col0 + l] = v[l];
              if (cq == 0) {  // row -> (output row, z_base row)
                rowinfo[r] = valid ? (long long)(cell * S + sr) : -1ll;
                rowinfo[128 + r] = (long long)fi;
              }
              epi_sync();
              const bool add_zb = !cb.per_row;
              int r2 = tid / L, l2 = tid - r2 * L;
              const int dr = 512 / L, dl = 512 - dr * L;
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
              if (need_dist) epi_sync();   // zf / rowinfo consumed before the reference rows and the Gram operands overwrite them
            }
            if (need_dist) {
              // Distances are invariant under a per-cell shift: every row is re-centred on the cell's sample 0 (keeps the
              // norms, and with them the cancellation in n_i + n_j - 2 g, small; the same choice as k_dist_gram_tc).
              float* cref_s = reinterpret_cast<float*>(slot + sm.off_cref);
              uint8_t* zh = slot + sm.off_zh;
              uint8_t* zl = slot + sm.off_zl;
              if (row_in_tile && s == 0) {
#pragma unroll
                for (int l = 0; l < 16; ++l)
                  if (l < ncol) cref_s[lc * Kz + col0 + l] = v[l];
              }
              epi_sync();
              const float* crow = cref_s + (row_in_tile ? lc : 0) * Kz + col0;
              float nrm_p = 0.f;
#pragma unroll
              for (int cc = 0; cc < 2; ++cc) {
                if (cc * 8 < ncol) {
                  const int c = (col0 >> 3) + cc;
                  uint32_t hw[4], lw[4];
#pragma unroll
                  for (int j2 = 0; j2 < 4; ++j2) {
                    const int k0 = cc * 8 + 2 * j2;
                    const float2 cr = *reinterpret_cast<const float2*>(crow + k0);
                    const float a = v[k0] - cr.x;
                    const float b = v[k0 + 1] - cr.y;
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
              const float2 st = f4_exchange(xch, xb, cq, r, q, make_float2(nrm_p, 0.f));
              xb ^= 1;
              if (cq == 0) reinterpret_cast<float*>(smem + sm.nrm[0] + sl * 512)[r] = st.x;
              fence_proxy_async();
            }
          } else {
            // ================= distance epilogue: this thread's 32-column quarter of Gram row r =================
            const int64_t cell = valid ? (gp.order ? (int64_t)gp.order[pos] : pos) : 0;
            const float* nrm_s = reinterpret_cast<const float*>(smem + sm.nrm[0] + sl * 512);
            uint8_t* zh = slot + sm.off_zh;
            uint8_t* zl = slot + sm.off_zl;
            // group-run bookkeeping (cells arrive sorted by composite code: a tile is uniform iff its end points agree)
            bool use_acc = false;
            if (gp.n_keys > 0) {
              const int64_t p0 = tile * cells_per_tile;
              const int64_t p1 = min(p0 + cells_per_tile, n_cells) - 1;
              const int first_code = gp.comp_sorted[min(p0, n_cells - 1)];
              use_acc = first_code == gp.comp_sorted[p1];
              if (use_acc) {
                if (chunk_needed && cur_valid && cur_code != first_code) f4_flush_acc(t_acc_lane, cq, gp, cur_cell, S, c0, row_in_tile, r);
                cur_code = first_code; cur_valid = true; cur_cell = gp.order[p0];
              }
            }
            if (chunk_needed) {
              const int c = cq;
              const float n_i = nrm_s[r];
              float g[32];
              tmem_ld32(t_lane + c * 32, g);
              tmem_ld_wait();
              // fast path (warp-uniform): the whole chunk lies inside every lane's cell and every row is valid -- no bounds checks;
              // the diagonal can only sit in the chunk of this warp's own rows (entry e == lane); cancellation-dominated entries
              // are only DETECTED here (min over the entries of d^2 - 2^-12 (n_i + n_j)) and located when there is one (rare)
              const bool full = __all_sync(0xffffffffu, valid && c * 32 >= c0 && c * 32 + 32 <= c0 + S);
              if (full) {
                const float tmin = (c == q) ? dist_chunk_fast<true>(g, nrm_s + c * 32, n_i, lane)
                                            : dist_chunk_fast<false>(g, nrm_s + c * 32, n_i, lane);
                if (__any_sync(0xffffffffu, tmin < 0.f)) {  // rare: near-duplicate samples -> exact direct-difference distance
                  uint32_t flagged = 0u;
                  if (tmin < 0.f) {
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                      if (c * 32 + e != r && g[e] * g[e] < (n_i + nrm_s[c * 32 + e]) * (1.0f / 4096.0f)) flagged |= 1u << e;
                  }
                  if (__any_sync(0xffffffffu, flagged != 0u)) exact_fix_warp<2048u>(g, flagged, zh, zl, r, c * 32, L);
                }
              } else {
                const float2 ni2 = make_float2(n_i, n_i), m2 = make_float2(-2.0f, -2.0f), thr2 = make_float2(-1.0f / 4096.0f, -1.0f / 4096.0f);
                uint32_t flagged = 0u;
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                  const float4 nj = *reinterpret_cast<const float4*>(nrm_s + c * 32 + 4 * e4);
                  const float2 njp[2] = {make_float2(nj.x, nj.y), make_float2(nj.z, nj.w)};
#pragma unroll
                  for (int e2 = 0; e2 < 2; ++e2) {
                    const int e = 4 * e4 + 2 * e2, j = c * 32 + e;     // columns j, j+1 = tile rows of the other samples
                    const float2 sn = __fadd2_rn(ni2, njp[e2]);
                    const float2 d2 = __ffma2_rn(make_float2(g[e], g[e + 1]), m2, sn);
                    const float2 tt = __ffma2_rn(sn, thr2, d2);        // < 0: cancellation dominated
                    const bool in0 = valid && (unsigned)(j - c0) < (unsigned)S && j != r;
                    const bool in1 = valid && (unsigned)(j + 1 - c0) < (unsigned)S && j + 1 != r;
                    if (in0 && tt.x < 0.f) flagged |= 1u << e;
                    if (in1 && tt.y < 0.f) flagged |= 2u << e;
                    g[e] = in0 ? sqrt_approx(fmaxf(d2.x, 0.f)) : 0.f;
                    g[e + 1] = in1 ? sqrt_approx(fmaxf(d2.y, 0.f)) : 0.f;
                  }
                }
                if (__any_sync(0xffffffffu, flagged != 0u))  // rare: near-duplicate samples -> exact direct-difference distance
                  exact_fix_warp<2048u>(g, flagged, zh, zl, r, c * 32, L);
              }
              if (cell_out != nullptr && (S & 3) == 0) {
                // Rows go through a per-warp shared-memory tile so that every store instruction writes full 32-byte sectors
                // (16 rows x 32 bytes) instead of 16 bytes of 32 different rows.  Four 8-column sub-steps; 16-byte units
                // XOR-swizzled with the row: no bank conflicts on either side.  The rows of a warp may belong to different
                // cells: the destination row and the cell's first column travel by shuffle.
                float4* T = reinterpret_cast<float4*>(slot + sm.off_T) + warp * 64;
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
                for (int e = 0; e < 32; ++e) {
                  const int j = c * 32 + e - c0;
                  if (j >= 0 && j < S) dst[j] = g[e];
                }
              }
              if (use_acc) {
                float acc[32];
                tmem_ld32(t_acc_lane + c * 32, acc);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                  const float2 t = __fadd2_rn(make_float2(acc[2 * e], acc[2 * e + 1]), make_float2(g[2 * e], g[2 * e + 1]));
                  acc[2 * e] = t.x; acc[2 * e + 1] = t.y;
                }
                tmem_st32(t_acc_lane + c * 32, reinterpret_cast<uint32_t*>(acc));
                tmem_st_wait();
              } else if (gp.n_keys > 0 && valid) {  // mixed tile: straight to the float64 sums
                for (int k = 0; k < gp.n_keys; ++k) {
                  const int gk = gp.codes[k][cell];
                  if (gk < 0 || gk >= gp.n_groups[k]) continue;
                  double* base = gp.sums[k] + ((size_t)gk * S + s) * S;
#pragma unroll
                  for (int e = 0; e < 32; ++e) {
                    const int j = c * 32 + e - c0;
                    if (j >= 0 && j < S && j != s) atomicAdd(base + j, (double)g[e]);
                  }
                }
              }
            }
            tc_fence_before();
          }
          F4_STAMP(2);
          if (ph < nsteps) {   // this warp's part of the slot's next GEMM operands is in place
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctrl->ready[sl]);
          }
        }
      }
    }
    if (gp.n_keys > 0 && cur_valid && chunk_needed) f4_flush_acc(t_acc_lane, cq, gp, cur_cell, S, c0, row_in_tile, r);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// fused distances: 4 <= S <= 128, l2; chain-only (z output): any S / n_latent the chain supports
inline bool tc_fused4_supported(const TcNet& tc, int S, int norm) {
  return tc.ready && tc.fused4_smem_bytes != 0 && norm == MRVI_NORM_L2 && S >= 4 && S <= 128;
}

inline int tc_fused4_prepare(TcNet& tc, int L) {
  const size_t bytes = f4_smem_map(tc.wimg_bytes, tc.K1, tc.K3, L).total + 1024;
  if (bytes > (size_t)kMaxSmemOptin) return 0;  // unavailable for this shape: the second version / unfused kernels are used
  if (cudaFuncSetAttribute(k_chain_fused4_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin) != cudaSuccess) {
    g_tc_error = "cudaFuncSetAttribute(k_chain_fused4_tc) failed";
    return MRVI_ERR_CUDA;
  }
  tc.fused4_smem_bytes = bytes;
  return 0;
}

inline void tc_launch_fused4(const TcNet& tc, const NetW& net, int64_t n, const CellBuf& cb, float* z_out, float* cell_out,
                             const GroupPlan& gp, cudaStream_t st) {
  const int S = net.S;
  int cpt = 1, tpc = 1;
  int64_t n_tiles;
  if (S <= kTcRows) { cpt = kTcRows / S; n_tiles = (n + cpt - 1) / cpt; }
  else { tpc = (S + kTcRows - 1) / kTcRows; n_tiles = n * tpc; }   // chain only
  if (n_tiles == 0) return;
  const int grid = (int)std::min<int64_t>((n_tiles + 1) / 2, tc.sm_count);
  const int64_t tiles_per_cta = (n_tiles + grid - 1) / grid;
  k_chain_fused4_tc<<<grid, kF4Threads, tc.fused4_smem_bytes, st>>>(tc, cb, n, S, net.L, net.Lq, z_out, cell_out, gp, n_tiles, cpt,
                                                                    tiles_per_cta, tpc);
}

}  // namespace mrvi
