// Tensor-core pairwise distances for 128 < S <= 256 (BASELINE config 5: "large-S Gram stress").
//
// The fused chain kernel (tc_fused.cuh) holds one 128-row tile per warpgroup; a cell with more than 128
// samples does not fit it, so the chain runs unfused (k_chain_tc writes z, 4 S L bytes per cell) and this
// kernel forms the S x S matrix from z on the tensor cores instead of the fp32 direct-difference kernel:
//
//   * one persistent CTA per SM, 512 threads; a CTA owns a contiguous range of SORTED cell positions;
//   * prep (all threads, two per sample row): z rows are re-centred on the mean of four of the cell's rows (any per-cell
//     shift cancels in the distances; it keeps the norms small), split into (hi, lo) f16 and stored as a
//     K-major UMMA operand of 256 rows, double buffered so that the prep of cell p+1 overlaps the MMAs of p;
//   * the matrix is covered by the three 128 x 128 blocks (0,0), (0,1), (1,1) -- D is symmetric, block (1,0)
//     is written / accumulated as the transpose of (0,1).  Warpgroup g owns columns [64 g, 64 g + 64) of
//     every block: it issues its own N = 64 MMAs (Zh Zh^T + Zh Zl^T + Zl Zh^T, fp32 accumulate in its half of
//     a 128-column TMEM work region), waits on its own mbarrier and runs the epilogue with two threads per
//     row, so the two warpgroups overlap each other's MMA latency without any cross-warpgroup barrier;
//   * epilogue: d = sqrt(n_i + n_j - 2 g), exact direct-difference recompute when the squared distance is
//     cancellation dominated (< 2^-12 (n_i + n_j)), diagonal forced to zero; rows go to cell_out and / or
//     into three 128-column fp32 TMEM accumulators that live across the cells of a run of equal composite
//     group code and are flushed with float64 atomics when the code changes (each warpgroup flushes its own
//     columns; no lock).
// TMEM: [0,128) work, [128,512) accumulators of the three blocks.
#pragma once
#include "tc_encoder.cuh"
#include "tc_fused.cuh"

namespace mrvi {

#ifdef MRVI_DBG_DIST
// clock64 timeline of CTA 0 (tools/dbg): [cell][slot]
__device__ long long g_dbg_dist[64 * 32];
#define DBG_STAMP(cellidx, slot) do { if (blockIdx.x == 0 && (cellidx) < 64) g_dbg_dist[(cellidx) * 32 + (slot)] = clock64(); } while (0)
#else
#define DBG_STAMP(cellidx, slot) do { } while (0)
#endif

constexpr int kDistTcThreads = 576;   // 16 epilogue warps (2 warpgroups) + 2 issuer warps (one per warpgroup: MMAs; the first also the bulk copies)
constexpr int kDistTcWorkers = 512;

// Gram entries with d^2 < 2^-12 (n_i + n_j) are recomputed by direct differences (the same threshold as the fused kernel).
// G[i][j] and G[j][i] are the same products accumulated in a different order, a few ulp of G apart: just above the threshold
// that is |d_ij - d_ji| up to ~2.5e-4 d on a handful of the 10^9 entries of a chunk (tests/test_gpu_scale.py bounds it by
// 5e-4, half the mode's accuracy bar; MRVI_PRECISION_FP32 is exactly symmetric).  A 2^-9 threshold removes it but was
// measured at +30 % on the fused kernel (C4 7.8 -> 10.3 ms: the exact path is warp-serial), so it is not used.
constexpr int kDistThrLog2 = 12;
constexpr float kDistThr = 1.0f / (float)(1 << kDistThrLog2);
constexpr uint32_t kDistLbo = 4096 + 16;  // bytes between K chunks: 256 rows x 16 B + one row of padding, which makes the
                                          // prep's 16-byte stores (8 chunks x 4 rows per warp) bank-conflict free

struct DistTcSmem {
  uint32_t z[2];      // per buffer: zh | zl, each 256 rows x Kz f16, K-major, LBO = kDistLbo
  uint32_t part[2];   // per buffer: [256] fp32 squared norms of the (hi + lo) operand rows
  uint32_t stage;     // raw fp32 z of the next cell ([S][L], landed by one bulk copy)
  uint32_t tr;        // store transposition tiles: 16 warps x [32 rows x 16 columns] fp32 (0 bytes when it does not fit)
  uint32_t tr_bytes;
  uint32_t ctrl;
  uint32_t total;
  int Kz;
};

__host__ __device__ inline DistTcSmem dist_tc_smem(int L, int S = 256) {
  DistTcSmem m;
  m.Kz = L <= 32 ? 32 : 64;
  const uint32_t zb = (2u * (uint32_t)(m.Kz >> 3) * kDistLbo + 127u) & ~127u;
  uint32_t o = 0;
  m.z[0] = o; o += zb;
  m.z[1] = o; o += zb;
  m.part[0] = o; o += 2048;
  m.part[1] = o; o += 2048;
  m.stage = o; o += (((uint32_t)S * (uint32_t)L * 4u) + 127u) & ~127u;
  m.tr = o;
  m.tr_bytes = (o + 16u * 2048u + 64u <= 227u * 1024u) ? 16u * 2048u : 0u;
  o += m.tr_bytes;
  m.ctrl = o; o += 64;
  m.total = o;
  return m;
}

// KEEP: per-cell output requested (compiled out of the groupby-only instantiation: its store path costs registers)
template <bool KEEP>
__global__ void __launch_bounds__(kDistTcThreads, 1)
k_dist_gram_tc(const float* __restrict__ z, int64_t n_cells, int S, int L, float* __restrict__ cell_out_arg, GroupPlan gp,
               int64_t cells_per_cta) {
  float* const cell_out = KEEP ? cell_out_arg : nullptr;
  extern __shared__ __align__(1024) uint8_t smem_tc[];
  uint8_t* const smem = smem_tc;
  const DistTcSmem sm = dist_tc_smem(L, S);
  const int Kz = sm.Kz;
  uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + sm.ctrl);       // [2], one per warpgroup
  uint64_t* bar_z = bar_mma + 2;                                         // raw z of the next cell landed
  uint64_t* bar_free = bar_mma + 3;                                      // [2] work columns of warpgroup g drained
  uint64_t* bar_ops = bar_mma + 5;                                       // operands of the next cell written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + sm.ctrl + 48);
  float* stage = reinterpret_cast<float*>(smem + sm.stage);
  const uint32_t zcell_bytes = (uint32_t)S * (uint32_t)L * 4u;
  // one bulk copy per cell needs 16-byte aligned, 16-byte multiple blocks; otherwise a cooperative copy
  const bool use_tma = (zcell_bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(z) & 15) == 0;
  uint32_t zphase = 0;
  const int tid = threadIdx.x, wg = tid >> 8, wt = tid & 255, sub = wt >> 7, r = wt & 127, warp_q = (wt >> 5) & 3;
  if (tid == 0) {
    mbar_init(&bar_mma[0], 1);
    mbar_init(&bar_mma[1], 1);
    mbar_init(bar_z, 1);
    mbar_init(&bar_free[0], 256);
    mbar_init(&bar_free[1], 256);
    mbar_init(bar_ops, kDistTcWorkers);
    fence_mbar_init();
  }
  if (tid < 32) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_off = (uint32_t)(warp_q * 32) << 16;
  const int colw = wg * 64 + sub * 32;                       // this thread's 32 columns inside a block
  const uint32_t t_work = tmem_base + lane_off + colw;
  const uint32_t t_acc0 = tmem_base + 128 + lane_off + colw;  // + 128 * block
  const uint32_t idesc64 = umma_idesc_f16(64);
  const uint32_t zbytes = (uint32_t)(Kz >> 3) * kDistLbo;     // one of (zh, zl)
  const int nblk = S > 128 ? 3 : 1;                           // S <= 128: the single block (0,0) (stage 3 alone on given z)
  uint32_t phase = 0;

  const int64_t p0 = (int64_t)blockIdx.x * cells_per_cta;
  const int64_t p1 = min(n_cells, p0 + cells_per_cta);
  if (p0 >= p1) {  // nothing to do (still release TMEM)
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem_base, 512);
    return;
  }
  const bool worker = tid < kDistTcWorkers;
  if (gp.n_keys > 0 && worker) {  // clear this thread's accumulator columns
    uint32_t zz[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) zz[q] = 0u;
    for (int b = 0; b < nblk; ++b) tmem_st32(t_acc0 + b * 128, zz);
    tmem_st_wait();
  }

  // ---- raw z of the cell at sorted position p -> staging: one bulk copy by the issuer thread, or (unaligned
  //      shapes) a cooperative coalesced copy by the workers right before the prep
  auto cell_src = [&](int64_t p) { return z + (gp.order ? (int64_t)gp.order[p] : p) * S * (int64_t)L; };
  auto workers_sync = [&]() { asm volatile("bar.sync 3, 512;" ::: "memory"); };
  auto stage_ready = [&](int64_t p) {  // workers: staging holds cell p
    if (use_tma) { mbar_wait(bar_z, zphase); zphase ^= 1; }
    else {
      const float* src = cell_src(p);
      for (int idx = tid; idx < S * L; idx += kDistTcWorkers) stage[idx] = __ldg(src + idx);
      workers_sync();
    }
  };
  // ---- operand prep of the staged cell into buffer buf (all 512 threads).  The nck = Kz / 8 chunks of a row sit in
  // adjacent lanes: a thread keeps its chunk (so the reference row's chunk stays in registers), the row norm is a
  // 2-3 step shuffle reduction, and with the padded chunk stride the 16-byte operand stores are conflict free.
  auto prep = [&](int buf) {
    const int nck = Kz >> 3, lg = nck == 8 ? 3 : 2;
    const int c = tid & (nck - 1), rr = tid >> lg, rows_per_pass = kDistTcWorkers >> lg;
    uint8_t* zh = smem + sm.z[buf];
    uint8_t* zl = zh + zbytes;
    float* nrm = reinterpret_cast<float*>(smem + sm.part[buf]);
    float2 z0p[4];
#pragma unroll
    for (int j2 = 0; j2 < 4; ++j2) {
      const int k0 = c * 8 + 2 * j2;
      // centre = mean of four of the cell's rows (samples 0, S/4, S/2, 3S/4): closer to the cell's mean than one sample, so
      // fewer entries fall under the exact-recompute threshold (see the fused kernel; the whole tile is staged here, so
      // four rows cost four loads per chunk element and cell)
      float cx = 0.f, cy = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float* zq = stage + ((q * S) >> 2) * L;
        cx += k0 < L ? zq[k0] : 0.f;
        cy += k0 + 1 < L ? zq[k0 + 1] : 0.f;
      }
      z0p[j2] = make_float2(-0.25f * cx, -0.25f * cy);
    }
    const bool even = (L & 1) == 0;
#pragma unroll 2
    for (int row = rr; row < (nblk == 3 ? 256 : 128); row += rows_per_pass) {
      const bool vrow = row < S;
      const float* zr = stage + (vrow ? row : 0) * L + c * 8;
      uint32_t hw[4], lw[4];
      float2 n2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) {
        const int k0 = c * 8 + 2 * j2;
        float2 a;
        if (even) a = (vrow && k0 < L) ? *reinterpret_cast<const float2*>(zr + 2 * j2) : make_float2(0.f, 0.f);  // L even: k0 + 1 < L too
        else a = make_float2((vrow && k0 < L) ? zr[2 * j2] : 0.f, (vrow && k0 + 1 < L) ? zr[2 * j2 + 1] : 0.f);
        a = __fadd2_rn(a, vrow ? z0p[j2] : make_float2(0.f, 0.f));
        split2v(a, hw[j2], lw[j2]);
        n2 = __ffma2_rn(a, a, n2);  // (hi + lo differs from a by 2^-22 relative: the level of the omitted lo.lo term)
      }
      *reinterpret_cast<uint4*>(zh + c * kDistLbo + row * 16) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(zl + c * kDistLbo + row * 16) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      float nr = n2.x + n2.y;
      nr += __shfl_xor_sync(0xffffffffu, nr, 1);
      nr += __shfl_xor_sync(0xffffffffu, nr, 2);
      if (nck == 8) nr += __shfl_xor_sync(0xffffffffu, nr, 4);
      if (c == 0) nrm[row] = nr;
    }
  };

  // ---- this warpgroup's N = 64 slice of Gram block (bi, bj) of buffer buf
  // (descriptor of address a + off = descriptor of a + (off >> 4): the address field holds a >> 4 and cannot carry
  //  out of its 14 bits for shared-memory addresses)
  const uint64_t dbase0 = umma_desc(smem_u32(smem + sm.z[0]), kDistLbo, 128), dbase1 = umma_desc(smem_u32(smem + sm.z[1]), kDistLbo, 128);
  auto issue_block = [&](int g_, int buf, int bi, int bj, uint32_t tb) {
    // only the 14-bit address field (low word) differs between the descriptors
    const uint32_t hi = (uint32_t)(dbase0 >> 32);
    const uint32_t lo_h = (uint32_t)(buf ? dbase1 : dbase0), lo_l = lo_h + (zbytes >> 4);
    const uint32_t a16 = (uint32_t)bi * 128u, b16 = (uint32_t)bj * 128u + (uint32_t)g_ * 64u;  // row offsets, 16-byte units
    uint32_t ah = lo_h + a16, al = lo_l + a16, bh = lo_h + b16, bl = lo_l + b16;
    const uint32_t d = tb + g_ * 64;
    auto mk = [hi](uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
    // cross terms first, the hi.hi term last: entries (i, j) and (j, i) of the diagonal blocks then differ by one ulp in rare
    // cases only (see the Gram issue of tc_fused.cuh)
    const uint32_t ah0 = ah, bh0 = bh;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (ks < Kz / 16) {
        umma_f16(d, mk(ah), mk(bl), idesc64, ks == 0 ? 0u : 1u);
        umma_f16(d, mk(al), mk(bh), idesc64, 1u);
        ah += 2u * (kDistLbo >> 4); al += 2u * (kDistLbo >> 4); bh += 2u * (kDistLbo >> 4); bl += 2u * (kDistLbo >> 4);  // two K chunks per 16-wide slice
      }
    }
    ah = ah0; bh = bh0;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      if (ks < Kz / 16) {
        umma_f16(d, mk(ah), mk(bh), idesc64, 1u);
        ah += 2u * (kDistLbo >> 4); bh += 2u * (kDistLbo >> 4);
      }
    }
    umma_commit(&bar_mma[g_]);
  };

  // ---- flush this thread's accumulator columns of the three blocks into the float64 group sums
  auto flush = [&](int64_t rep_cell) {
#pragma unroll 1
    for (int b = 0; b < nblk; ++b) {
      const int bi = b == 2 ? 1 : 0, bj = b == 0 ? 0 : 1;
      float a[32];
      tmem_ld32(t_acc0 + b * 128, a);
      tmem_ld_wait();
      const int i = bi * 128 + r;
      if (i < S) {
#pragma unroll 1
        for (int k = 0; k < gp.n_keys; ++k) {
          const int g = gp.codes[k][rep_cell];
          if (g < 0 || g >= gp.n_groups[k]) continue;
          double* base = gp.sums[k] + (size_t)g * S * S;
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int j = bj * 128 + colw + q;
            if (j < S && j != i) {
              atomicAdd(base + (size_t)i * S + j, (double)a[q]);
              if (b == 1) atomicAdd(base + (size_t)j * S + i, (double)a[q]);
            }
          }
        }
      }
      uint32_t zz[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) zz[q] = 0u;
      tmem_st32(t_acc0 + b * 128, zz);
    }
    tmem_st_wait();
  };

  if (!worker) {
    // ================= issuer warps: warp 16 -> warpgroup 0 + the bulk copies of z, warp 17 -> warpgroup 1.  The branch is
    // warp-uniform (warp index through a shuffle) with elect.sync around the MMAs / copies: issued from `if (lane == 0)`
    // code every tcgen05.mma costs ~92 clk instead of 42-61 (see elect_one in tc_kernels.cuh)
    {
      const int g_ = __shfl_sync(0xffffffffu, (tid - kDistTcWorkers) >> 5, 0);
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      auto tma = [&](int64_t p) {
        if (!use_tma || g_ != 0) return;
        if (elect_one()) {
          mbar_expect_tx(bar_z, zcell_bytes);
          bulk_g2s(stage, cell_src(p), zcell_bytes, bar_z);
        }
        __syncwarp();
      };
      uint32_t ops_phase = 0, free_phase = 1;  // a fresh mbarrier passes a wait on parity 1
      tma(p0);
      mbar_wait(bar_ops, ops_phase); ops_phase ^= 1;    // operands of p0 written, staging consumed
      if (p0 + 1 < p1) tma(p0 + 1);
      for (int64_t p = p0; p < p1; ++p) {
        const int buf = (int)((p - p0) & 1);
#pragma unroll 1
        for (int b = 0; b < nblk; ++b) {
          mbar_wait(&bar_free[g_], free_phase); free_phase ^= 1;
          tc_fence_after();
          if (g_ == 0) DBG_STAMP(p - p0, 0 + 2 * b);
          if (elect_one()) issue_block(g_, buf, b == 2 ? 1 : 0, b == 0 ? 0 : 1, tb);
          __syncwarp();
          if (g_ == 0) DBG_STAMP(p - p0, 1 + 2 * b);
        }
        if (p + 1 < p1) {
          mbar_wait(bar_ops, ops_phase); ops_phase ^= 1;  // prep(p+1) done by every worker
          if (p + 2 < p1) tma(p + 2);                     // lands while the workers run the epilogues of p
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    return;
  }

  // ================= workers =================
  stage_ready(p0);
  prep(0);
  fence_proxy_async();
  mbar_arrive(bar_ops);
  workers_sync();

  int cur_code = -1;
  int64_t cur_cell = -1;
  for (int64_t p = p0; p < p1; ++p) {
    const int buf = (int)((p - p0) & 1);
    if (tid == 0) DBG_STAMP(p - p0, 22);
    const int64_t cell = gp.order ? (int64_t)gp.order[p] : p;
    const int code = gp.n_keys ? gp.comp_sorted[p] : 0;
    if (tid == 0) DBG_STAMP(p - p0, 8);
    if (p + 1 < p1) {  // overlaps the MMAs of block (0,0), which the issuer warp has already queued
      stage_ready(p + 1);
      if (tid == 0) DBG_STAMP(p - p0, 9);
      prep(buf ^ 1);
      fence_proxy_async();
      mbar_arrive(bar_ops);
    }

    if (tid == 0) DBG_STAMP(p - p0, 10);
#ifdef MRVI_DBG_DIST
    const long long t_prep_done = clock64();
#endif
    if (gp.n_keys && cur_cell >= 0 && code != cur_code) flush(cur_cell);
    if (tid == 0) DBG_STAMP(p - p0, 23);
    cur_code = code; cur_cell = cell;
    const uint8_t* zh = smem + sm.z[buf];
    const uint8_t* zl = zh + zbytes;
    const float* part = reinterpret_cast<const float*>(smem + sm.part[buf]);
#pragma unroll 1
    for (int b = 0; b < nblk; ++b) {
      const int bi = b == 2 ? 1 : 0, bj = b == 0 ? 0 : 1;
      if (tid == 0) DBG_STAMP(p - p0, 11 + 3 * b);
      mbar_wait(&bar_mma[wg], phase); phase ^= 1;
      tc_fence_after();
      if (tid == 0) DBG_STAMP(p - p0, 12 + 3 * b);
      float g[32];
      tmem_ld32(t_work, g);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&bar_free[wg]);  // this thread has read its work columns: the issuer may queue the next block
      if (tid == 0) DBG_STAMP(p - p0, 13 + 3 * b);
      const int i = bi * 128 + r, jbase = bj * 128 + colw;
      const float n_i = part[i];
      // warp-uniform cases: `full` = every row of the warp and the whole chunk lie inside the S x S matrix;
      // `diag_chunk` = the chunk holds the diagonal entries of this warp's rows (diagonal blocks only)
      const bool full = __all_sync(0xffffffffu, i < S) && jbase + 32 <= S;
      const bool diag_chunk = (b != 1) && ((colw >> 5) == warp_q);
      if (full) {
        // fast path: no bounds checks; cancellation-dominated entries are only DETECTED here
        const float tmin = diag_chunk ? dist_chunk_fast<true, kDistThrLog2>(g, part + jbase, n_i, tid & 31)
                                      : dist_chunk_fast<false, kDistThrLog2>(g, part + jbase, n_i, tid & 31);
        if (__any_sync(0xffffffffu, tmin < 0.f)) {  // rare: near-duplicate samples -> exact direct-difference distance
          uint32_t flagged = 0u;
          if (tmin < 0.f) {
#pragma unroll
            for (int q = 0; q < 32; ++q)
              if (jbase + q != i && g[q] * g[q] < (n_i + part[jbase + q]) * kDistThr) flagged |= 1u << q;
          }
          exact_fix_warp<kDistLbo>(g, flagged, zh, zl, i, jbase, L);
        }
      } else {
        const float2 ni2 = make_float2(n_i, n_i), m2 = make_float2(-2.0f, -2.0f), thr2 = make_float2(-kDistThr, -kDistThr);
        uint32_t flagged = 0u;
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 nj = *reinterpret_cast<const float4*>(part + jbase + 4 * q4);
          const float2 njp[2] = {make_float2(nj.x, nj.y), make_float2(nj.z, nj.w)};
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int q = 4 * q4 + 2 * e2, j = jbase + q;
            const float2 sn = __fadd2_rn(ni2, njp[e2]);
            const float2 d2 = __ffma2_rn(make_float2(g[q], g[q + 1]), m2, sn);
            const float2 tt = __ffma2_rn(sn, thr2, d2);
            const bool in0 = i < S && j < S && j != i;
            const bool in1 = i < S && j + 1 < S && j + 1 != i;
            if (in0 && tt.x < 0.f) flagged |= 1u << q;
            if (in1 && tt.y < 0.f) flagged |= 2u << q;
            g[q] = in0 ? sqrt_approx(fmaxf(d2.x, 0.f)) : 0.f;
            g[q + 1] = in1 ? sqrt_approx(fmaxf(d2.y, 0.f)) : 0.f;
          }
        }
        if (__any_sync(0xffffffffu, flagged != 0u)) exact_fix_warp<kDistLbo>(g, flagged, zh, zl, i, jbase, L);
      }
      if (cell_out != nullptr && (S & 3) == 0 && sm.tr_bytes != 0) {
        // rows through a per-warp shared-memory tile so that every store instruction writes full 32-byte sectors:
        // a thread-per-row float4 store touches 32 rows = 32 half-used sectors.  Two 16-column sub-steps; 16-byte
        // units are XOR-swizzled with the row so that neither side has bank conflicts.
        const int lane = tid & 31;
        float4* T = reinterpret_cast<float4*>(smem + sm.tr + (tid >> 5) * 2048);
        const int row0 = bi * 128 + warp_q * 32;   // first row of this warp
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            T[lane * 4 + (q ^ ((lane >> 1) & 3))] = make_float4(g[h2 * 16 + 4 * q], g[h2 * 16 + 4 * q + 1], g[h2 * 16 + 4 * q + 2], g[h2 * 16 + 4 * q + 3]);
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rl = it * 8 + (lane >> 2), unit = lane & 3;
            const int ii = row0 + rl, jj = jbase + h2 * 16 + unit * 4;
            const float4 v = T[rl * 4 + (unit ^ ((rl >> 1) & 3))];
            if (ii < S && jj < S) *reinterpret_cast<float4*>(cell_out + ((size_t)cell * S + ii) * S + jj) = v;
          }
          __syncwarp();
        }
      } else if (cell_out != nullptr && i < S) {
        float* dst = cell_out + ((size_t)cell * S + i) * S + jbase;
        if ((S & 3) == 0) {
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4)
            if (jbase + 4 * q4 < S)
              *reinterpret_cast<float4*>(dst + 4 * q4) = make_float4(g[4 * q4], g[4 * q4 + 1], g[4 * q4 + 2], g[4 * q4 + 3]);
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) if (jbase + q < S) dst[q] = g[q];
        }
      }
      if (cell_out != nullptr && i < S) {
        if (b == 1) {  // block (1,0) = transpose of (0,1): lanes hold consecutive i -> coalesced 128-byte rows
          float* dt = cell_out + (size_t)cell * S * S + i;
#pragma unroll
          for (int q = 0; q < 32; ++q) if (jbase + q < S) dt[(size_t)(jbase + q) * S] = g[q];
        }
      }
      if (gp.n_keys > 0) {
#pragma unroll
        for (int hq = 0; hq < 2; ++hq) {  // two 16-column halves: bounds the live registers
          float acc[16];
          tmem_ld16(t_acc0 + b * 128 + hq * 16, acc);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float2 t = __fadd2_rn(make_float2(acc[2 * q], acc[2 * q + 1]), make_float2(g[hq * 16 + 2 * q], g[hq * 16 + 2 * q + 1]));
            acc[2 * q] = t.x; acc[2 * q + 1] = t.y;
          }
          tmem_st16(t_acc0 + b * 128 + hq * 16, reinterpret_cast<uint32_t*>(acc));
        }
        tmem_st_wait();
      }
    }
    // the exact-recompute route and the norms read buffer buf: every worker is done with cell p before the prep
    // of cell p+2 overwrites it
    if (tid == 0) DBG_STAMP(p - p0, 20);
#ifdef MRVI_DBG_DIST
    if (blockIdx.x == 0 && (tid & 31) == 0 && (p - p0) < 32) g_dbg_dist[1024 + (p - p0) * 16 + (tid >> 5)] = clock64();
    if (blockIdx.x == 0 && (tid & 31) == 0 && (p - p0) < 24) g_dbg_dist[1536 + (p - p0) * 16 + (tid >> 5)] = t_prep_done;
#endif
    workers_sync();
    if (tid == 0) DBG_STAMP(p - p0, 21);
  }
  if (gp.n_keys && cur_cell >= 0) flush(cur_cell);
  tc_fence_before();
  __syncthreads();  // joins the issuer warp
  if (tid < 32) tmem_dealloc(tmem_base, 512);
}

// 128 < S <= 256: the large-S route of the path.  32 < S <= 128: only reached through mrvi_distances_device (stage 3
// alone on caller-provided z; the path itself runs the fused kernel there), one cell per iteration in block (0,0).
inline bool dist_tc_supported(int S, int L, int norm) { return norm == MRVI_NORM_L2 && S > 32 && S <= 256 && L <= 64; }

inline int dist_tc_prepare(int L, int S, size_t* smem_bytes) {
  const size_t bytes = dist_tc_smem(L, S).total;
  if (bytes > 227 * 1024) { *smem_bytes = 0; return 0; }
  if (cudaFuncSetAttribute(k_dist_gram_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin) != cudaSuccess ||
      cudaFuncSetAttribute(k_dist_gram_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin) != cudaSuccess) {
    g_tc_error = "cudaFuncSetAttribute(k_dist_gram_tc) failed";
    return MRVI_ERR_CUDA;
  }
  *smem_bytes = bytes;
  return 0;
}

inline void dist_tc_launch(const float* z, int64_t n, int S, int L, float* cell_out, const GroupPlan& gp, int sm_count,
                           size_t smem_bytes, cudaStream_t st) {
  if (n == 0) return;
  const int grid = (int)std::min<int64_t>(n, sm_count);
  const int64_t cpc = (n + grid - 1) / grid;
  if (cell_out)
    k_dist_gram_tc<true><<<(int)((n + cpc - 1) / cpc), kDistTcThreads, smem_bytes, st>>>(z, n, S, L, cell_out, gp, cpc);
  else
    k_dist_gram_tc<false><<<(int)((n + cpc - 1) / cpc), kDistTcThreads, smem_bytes, st>>>(z, n, S, L, nullptr, gp, cpc);
}

}  // namespace mrvi
