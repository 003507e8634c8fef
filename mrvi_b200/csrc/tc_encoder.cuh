// Whole encoder x -> u on the tensor cores (reference _module.py:173-202), one CTA = 128 cells.
//
// Phase 1 (HBM bound): H1 = log1p(X) . W0, the warp-specialised 3-stage pipeline of k_enc_gemm_tc.
// Phase 2 (the "tail", hidden behind the other resident CTA's phase 1): the 128-wide layers
//   h = gelu(gamma0[s] LN(H1 + b0) + beta0[s])                     ConditionalNormalization_0
//   h = gelu(gamma1[s] LN(Dense_1 h) + beta1[s]) + sample_effect[s]  ConditionalNormalization_1, Embed_0
//   per ResnetBlock:  t = relu(LN(Dense_0 h)) + h ;  h = relu(LN(Dense_out t))     (relu / relu)
//   u = Dense_mean h ;  then per row  LN(u), query tokens, z_base  (reference _module.py:142, 166-170,
//   _components.py:195-197)
// run as tcgen05 GEMMs whose A operand is the previous layer's activation kept IN TMEM as (hi, lo) f16
// pairs (TS-mode MMA, two 128-column regions used alternately; the residual `+ h` is re-read from
// the previous region) against (hi, lo) f16 weights streamed per layer by TMA into the freed
// pipeline stages:  D = Ah.Wh + Al.Wh + Ah.Wl  (~21 significant bits; single-f16 weights here would
// cost 4-8e-4 normwise on the distances).  Epilogues are thread-per-row (TMEM lane =
// cell), chunked by 32 columns, fp32.
#pragma once
#include <cuda.h>

#include "tc_kernels.cuh"

namespace mrvi {

constexpr int kEncMaxLayers = 1 + 2 * kMaxBlocks;
constexpr int kEncLayerBytes = 2 * 128 * 128 * 2;  // hi | lo of a [128 x 128] f16 operand
constexpr int kEncHeadMaxN = 96;                   // mean head [u | z_base] columns, padded to a multiple of 32

struct EncTail {
  const float *gamma0, *beta0, *gamma1, *beta1, *sample_effect;  // [S][128]
  const uint8_t* wl[kEncMaxLayers];  // Dense_1, then (Dense_0, Dense_out) per ResnetBlock
  const float* bl[kEncMaxLayers];    // biases [128]
  const float* lg[kEncMaxLayers];    // LayerNorm scale [128] (null for the conditional layer)
  const float* lb[kEncMaxLayers];
  int n_layers;                      // 1 + 2 * n_blocks
  const uint8_t* whead;              // head [Nh x 128] hi | lo:  h -> [u (Lq) | z_base (L)]  (z_base folded: W_mean . W_zb)
  const float* bhead;                // [Nh]  (b_mean ; b_mean . W_zb + b_zb)
  int Nh;                            // padded head width (32, 64 or 96)
  const float *uln_g, *uln_b;        // [Lq]
  const float* wq;                   // [Lq][16]   DenseGeneral_0 (query tokens)
  int has_zbase;                     // 0: isomorphic u/z (z_base = u)
  int Lq, L, S;
  int sm_count;
  bool ready;
};

__device__ __forceinline__ float f16lo(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))); }
__device__ __forceinline__ float f16hi(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }

__device__ __forceinline__ void tmem_ld16f(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void enc_tail_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// One tail epilogue, two threads per row: thread (row, half) owns columns [64 half, 64 half + 64) in four
// 16-column sub-chunks.   y = act(LN(D * scale + bias) * g + b) [+ add] [+ residual], written back in place
// as the next A operand: per 16-column chunk [8 hi words | 8 lo words] (f16 pairs), so K slice ks of the
// next MMA reads hi at column 16 ks and lo at 16 ks + 8.  LayerNorm partial sums are exchanged through
// shared memory.  ACT: 1 relu, 2 gelu.  g may be null.  t_res: TMEM address of the residual region or 0.
template <int ACT>
__device__ __forceinline__ void enc_ep(uint32_t t_d, uint32_t t_res, int half, int row, float2* xch, float scale,
                                       const float* __restrict__ bias, const float* __restrict__ g,
                                       const float* __restrict__ b, const float* __restrict__ add) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
  for (int p = 0; p < 4; ++p) {
    const int c0 = half * 64 + p * 16;
    float v[16];
    tmem_ld16f(t_d + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c0) + q);
      const float x0 = fmaf(v[4 * q], scale, b4.x), x1 = fmaf(v[4 * q + 1], scale, b4.y);
      const float x2 = fmaf(v[4 * q + 2], scale, b4.z), x3 = fmaf(v[4 * q + 3], scale, b4.w);
      s1 += (x0 + x1) + (x2 + x3);
      s2 = fmaf(x0, x0, s2); s2 = fmaf(x1, x1, s2); s2 = fmaf(x2, x2, s2); s2 = fmaf(x3, x3, s2);
    }
  }
  xch[half * 128 + row] = make_float2(s1, s2);
  enc_tail_sync();
  const float2 o = xch[(half ^ 1) * 128 + row];
  s1 += o.x; s2 += o.y;
  const float mean = s1 * (1.0f / 128);
  const float var = fmaxf(fmaf(s2, 1.0f / 128, -mean * mean), 0.f);
  const float rstd = 1.0f / sqrtf(var + kLnEps);
  const float nb = -mean * rstd;
#pragma unroll 1
  for (int p = 0; p < 4; ++p) {
    const int c0 = half * 64 + p * 16;
    float v[16];
    uint32_t rw[16];
    tmem_ld16f(t_d + c0, v);
    if (t_res) tmem_ld16f(t_res + c0, reinterpret_cast<float*>(rw));
    tmem_ld_wait();
    uint32_t out[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 bi = __ldg(reinterpret_cast<const float4*>(bias + c0) + q);
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(b + c0) + q);
      float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f), a4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (g) g4 = __ldg(reinterpret_cast<const float4*>(g + c0) + q);
      if (add) a4 = __ldg(reinterpret_cast<const float4*>(add + c0) + q);
      float y[4];
      const float xs[4] = {fmaf(v[4 * q], scale, bi.x), fmaf(v[4 * q + 1], scale, bi.y), fmaf(v[4 * q + 2], scale, bi.z),
                           fmaf(v[4 * q + 3], scale, bi.w)};
      const float gs[4] = {g4.x, g4.y, g4.z, g4.w}, bs[4] = {b4.x, b4.y, b4.z, b4.w}, as[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float t = fmaf(fmaf(xs[e], rstd, nb), gs[e], bs[e]);
        t = (ACT == 1) ? fmaxf(t, 0.f) : gelu_fast(t);
        y[e] = t + as[e];
      }
      if (t_res) {  // residual = previous activation (hi + lo): element 2j of the chunk in the low halves of words j / 8+j
        const uint32_t h0 = rw[2 * q], l0 = rw[8 + 2 * q], h1 = rw[2 * q + 1], l1 = rw[8 + 2 * q + 1];
        y[0] += f16lo(h0) + f16lo(l0); y[1] += f16hi(h0) + f16hi(l0);
        y[2] += f16lo(h1) + f16lo(l1); y[3] += f16hi(h1) + f16hi(l1);
      }
      split2(y[0], y[1], out[2 * q], out[8 + 2 * q]);
      split2(y[2], y[3], out[2 * q + 1], out[8 + 2 * q + 1]);
    }
    tmem_st16(t_d + c0, out);
  }
  tmem_st_wait();
}

// ------------------------------------------------------------------------------------------
// Persistent, warp-specialised encoder: ONE CTA per SM loops over 128-cell tiles.
//   warps 0-7   producers: stream X (coalesced LDG.128, two K blocks of register prefetch that runs across
//               tile boundaries), log1p, (hi, lo) split, K-major stores        -> HBM is busy all the time
//   warps 8-15  tail epilogues of the PREVIOUS tile (two threads per row), concurrently with the stream
//   warp 16     phase-1 MMA issuer (accumulator double-buffered in TMEM: ACC[tile & 1])
//   warp 17     tail MMA issuer (A operand from TMEM)          warp 18 / 19   TMA loaders (W0 blocks / layer weights)
// One CTA per SM also halves the number of concurrently streamed rows (DRAM row-buffer locality):
// measured 3.45 TB/s for phase 1 alone with 148 CTAs versus 1.4 TB/s effective with two CTAs per SM.
// TMEM: ACC0 [0,128)  ACC1 [128,256)  T0 [256,384)  T1 [384,512).
// ------------------------------------------------------------------------------------------
constexpr int kEncPStages = 4;
constexpr int kEncPThreads = 640;
constexpr size_t kEncPSmem = (size_t)kEncPStages * kEncStageBytes + kEncLayerBytes + 4096 + 1024;
// TMAX variant (16-byte aligned rows): X is streamed by the copy engine, not by the producers' registers.
//   raw X ring   kEncRX stages x [128 rows x 128 B]   ONE 2-D tensor copy (cp.async.bulk.tensor, 128 x 32 box of the
//                                                     [n, G] matrix, out-of-range rows / columns zero-filled) per K block,
//                                                     issued 4 K blocks ahead: the bytes in flight no longer depend on
//                                                     how fast the converters run.  (128 per-row bulk copies instead were
//                                                     measured at ~90 clk per copy per SM: 4x slower than the LDG path.)
//   W0 ring      kEncRW stages x 16 KiB (hi | lo)
//   A ring       kEncRA stages x (hi | lo) converted operand
// The producers become converters: raw stage (shared-memory latency instead of 1.5 us of HBM latency) -> log1p ->
// (hi, lo) split -> A stage.
constexpr int kEncRX = 4, kEncRW = 3, kEncRA = 2;
#ifdef MRVI_DBG_ENC
__device__ long long g_dbg_enc[64 * 16];  // clock64 timeline of CTA 0, per K block g < 64 (tools/dbg)
#define ENC_STAMP(gi, slot) do { if (blockIdx.x == 0 && (gi) >= 128 && (gi) < 192) g_dbg_enc[((gi) - 128) * 16 + (slot)] = clock64(); } while (0)
#else
#define ENC_STAMP(gi, slot) do { } while (0)
#endif
constexpr int kEncRawBytes = 128 * 128;
constexpr size_t kEncTSmem = (size_t)kEncRX * kEncRawBytes + (size_t)kEncRW * kEncBBytes + (size_t)kEncRA * 2 * kEncAHalf +
                             kEncLayerBytes + 4096 + 1024;

template <bool TMAX>
__global__ void __launch_bounds__(kEncPThreads, 1)
k_encoder_tc(const float* __restrict__ X, int64_t ldx, const int32_t* __restrict__ sample_idx, int64_t n_cells, int G,
             EncTc enc, const __grid_constant__ EncTail tail, CellBuf cb, int x_vec_ok, const __grid_constant__ CUtensorMap xmap) {
  // no-swizzle UMMA operands and bulk copies need 16-byte alignment only; keeping the plain array (no integer
  // re-alignment) lets the compiler see the shared address space (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem_tc[];
  uint8_t* const smem = smem_tc;
  // TMAX layout: [raw X ring | W0 ring | A ring | wbuf | barriers]
  uint8_t* const rawx = smem;
  uint8_t* const w0ring = smem + (size_t)kEncRX * kEncRawBytes;
  uint8_t* const aring = w0ring + (size_t)kEncRW * kEncBBytes;
  uint8_t* wbuf = TMAX ? aring + (size_t)kEncRA * 2 * kEncAHalf
                       : smem + (size_t)kEncPStages * kEncStageBytes;       // tail layer weights (hi | lo)
  uint64_t* bars = reinterpret_cast<uint64_t*>(wbuf + kEncLayerBytes);
  uint64_t* full = bars;                       // [stages] producers (256) + W0 expect_tx (1)   (TMAX: a_full[kEncRA], 256)
  uint64_t* empty = full + kEncPStages;        // [stages] tcgen05.commit                       (TMAX: a_empty[kEncRA])
  uint64_t* x_full = empty + kEncPStages;      // TMAX [kEncRX] raw rows landed (1 + tx)
  uint64_t* x_empty = x_full + kEncRX;         // TMAX [kEncRX] converters have the raw values in registers (256)
  uint64_t* w_full = x_empty + kEncRX;         // TMAX [kEncRW] W0 block landed (1 + tx)
  uint64_t* w_empty = w_full + kEncRW;         // TMAX [kEncRW] tcgen05.commit
  uint64_t* acc_full = w_empty + kEncRW;       // [2] phase-1 accumulator of a tile complete
  uint64_t* acc_free = acc_full + 2;           // [2] accumulator region may be overwritten
  uint64_t* bar_a = acc_free + 2;              // tail: next layer's activation is in TMEM (256 arrivals)
  uint64_t* bar_d = bar_a + 1;                 // tail: layer MMAs complete
  uint64_t* bar_w = bar_d + 1;                 // tail: layer weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 1);
  float2* xch = reinterpret_cast<float2*>(tmem_slot + 4);                   // [2][128] partial LayerNorm sums
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = enc.n_kblocks;
  const int NL = tail.n_layers;
  const int64_t n_tiles = (n_cells + 127) / 128;
  const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < kEncPStages; ++s) { mbar_init(full + s, TMAX ? 256 : 257); mbar_init(empty + s, 1); }
    for (int s = 0; s < kEncRX; ++s) { mbar_init(x_full + s, 1); mbar_init(x_empty + s, 256); }
    for (int s = 0; s < kEncRW; ++s) { mbar_init(w_full + s, 1); mbar_init(w_empty + s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(acc_full + b, 1); mbar_init(acc_free + b, 1); }
    mbar_init(bar_a, 256);
    mbar_init(bar_d, 1);
    mbar_init(bar_w, 1);
    fence_mbar_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_base_t = tmem_base;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);   // provably warp-uniform (MMA issue branches)

  if (TMAX && warp < 8) {
    // ------------------------------------------------------------ converters: raw X stage -> (hi, lo) A stage
    const int rsub = lane >> 2, chunk = lane & 3;
    const int64_t total = my_tiles * nkb;
    for (int64_t g = 0; g < total; ++g) {
      const int sx = (int)(g % kEncRX), sa = (int)(g % kEncRA);
      if (tid == 0) ENC_STAMP(g, 0);
      mbar_wait(x_full + sx, (uint32_t)((g / kEncRX) & 1));
      if (tid == 0) ENC_STAMP(g, 1);
      float4 v[4];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int row = warp * 16 + p * 8 + rsub;
        const float4* src = reinterpret_cast<const float4*>(rawx + (size_t)sx * kEncRawBytes + row * 128 + chunk * 32);
        v[2 * p] = src[0];
        v[2 * p + 1] = src[1];
      }
      uint32_t h[2][4], l[2][4];
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const float4 a = v[2 * p], b = v[2 * p + 1];
        split2v(log1p_fast2(make_float2(a.x, a.y)), h[p][0], l[p][0]);
        split2v(log1p_fast2(make_float2(a.z, a.w)), h[p][1], l[p][1]);
        split2v(log1p_fast2(make_float2(b.x, b.y)), h[p][2], l[p][2]);
        split2v(log1p_fast2(make_float2(b.z, b.w)), h[p][3], l[p][3]);
      }
      if (tid == 0) ENC_STAMP(g, 2);
      if (g >= kEncRA) mbar_wait(empty + sa, (uint32_t)((g / kEncRA - 1) & 1));
      if (tid == 0) ENC_STAMP(g, 4);
      uint8_t* stage = aring + (size_t)sa * 2 * kEncAHalf;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int row = warp * 16 + p * 8 + rsub;
        *reinterpret_cast<uint4*>(stage + chunk * kEncALbo + row * 16) = make_uint4(h[p][0], h[p][1], h[p][2], h[p][3]);
        *reinterpret_cast<uint4*>(stage + kEncAHalf + chunk * kEncALbo + row * 16) = make_uint4(l[p][0], l[p][1], l[p][2], l[p][3]);
      }
      // ONE proxy fence serves both directions: the stores above become visible to the MMAs (async proxy), and the
      // shared-memory loads of the raw stage (their values were consumed by the stores) are ordered before the copy
      // engine refills it.  Without it the refill raced the reads: a few rows per launch came out wrong.
      fence_proxy_async();
      mbar_arrive(full + sa);
      mbar_arrive(x_empty + sx);
      if (tid == 0) ENC_STAMP(g, 5);
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------ producers: flat loop over (tile, K block)
    const int rsub = lane >> 2, chunk = lane & 3;
    const int64_t total = my_tiles * nkb;
    float4 cur[4], nxt[4], nx2[4];
    auto load = [&](int64_t g, float4 (&dst)[4]) {
      const int64_t ti = g / nkb;
      const int kb = (int)(g - ti * nkb);
      const int64_t cell0 = ((int64_t)blockIdx.x + ti * gridDim.x) * 128;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int row = warp * 16 + p * 8 + rsub;
        const int64_t cell = cell0 + row;
        const int k = kb * kEncBK + chunk * 8;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (cell < n_cells) {
          const float* src = X + cell * ldx + k;
          if (x_vec_ok && k + 8 <= G) {
            a = __ldg(reinterpret_cast<const float4*>(src));
            b = __ldg(reinterpret_cast<const float4*>(src) + 1);
          } else {
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = (k + j < G) ? __ldg(src + j) : 0.f;
            a = make_float4(t[0], t[1], t[2], t[3]);
            b = make_float4(t[4], t[5], t[6], t[7]);
          }
        }
        dst[2 * p] = a;
        dst[2 * p + 1] = b;
      }
    };
    auto convert = [&](int64_t g, const float4 (&src)[4]) {
      const int s = (int)(g % kEncPStages);
      const int64_t use = g / kEncPStages;
      if (use > 0) mbar_wait(empty + s, (uint32_t)((use - 1) & 1));
      uint8_t* stage = smem + s * kEncStageBytes;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const int row = warp * 16 + p * 8 + rsub;
        const float4 a = src[2 * p], b = src[2 * p + 1];
        uint32_t h[4], l[4];
        split2(log1p_fast(a.x), log1p_fast(a.y), h[0], l[0]);
        split2(log1p_fast(a.z), log1p_fast(a.w), h[1], l[1]);
        split2(log1p_fast(b.x), log1p_fast(b.y), h[2], l[2]);
        split2(log1p_fast(b.z), log1p_fast(b.w), h[3], l[3]);
        *reinterpret_cast<uint4*>(stage + chunk * kEncALbo + row * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(stage + kEncAHalf + chunk * kEncALbo + row * 16) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      fence_proxy_async();
      mbar_arrive(full + s);
    };
    if (total > 0) load(0, cur);
    if (total > 1) load(1, nxt);
    for (int64_t g = 0; g < total; g += 3) {
      if (g + 2 < total) load(g + 2, nx2);
      convert(g, cur);
      if (g + 1 < total) {
        if (g + 3 < total) load(g + 3, cur);
        convert(g + 1, nxt);
      }
      if (g + 2 < total) {
        if (g + 4 < total) load(g + 4, nxt);
        convert(g + 2, nx2);
      }
    }
  } else if (warp < 16) {
    // ------------------------------------------------------------ tail epilogues: thread (row, half), row = TMEM lane
    const int et = tid - 256;
    const int row = et & 127, half = et >> 7;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    int64_t c = 0;  // running layer counter (phases of bar_a / bar_d / bar_w)
    for (int64_t ti = 0; ti < my_tiles; ++ti) {
      const int64_t cell = ((int64_t)blockIdx.x + ti * gridDim.x) * 128 + row;
      const bool valid = cell < n_cells;
      int sid = valid ? sample_idx[cell] : 0;
      sid = min(max(sid, 0), tail.S - 1);
      const uint32_t t_acc = tmem_base + (uint32_t)(ti & 1) * 128 + lane_off;
      const uint32_t t_T[2] = {tmem_base + 256 + lane_off, tmem_base + 384 + lane_off};
      mbar_wait(acc_full + (ti & 1), (uint32_t)((ti >> 1) & 1));
      tc_fence_after();
      // ConditionalNormalization_0 + gelu on H1 = D / 2^8 + b0          (reference _module.py:190-195)
      enc_ep<2>(t_acc, 0u, half, row, xch, 1.0f / kEncWScale, enc.b0, tail.gamma0 + (size_t)sid * 128,
                tail.beta0 + (size_t)sid * 128, nullptr);
      tc_fence_before();
      mbar_arrive(bar_a);
      for (int li = 0; li < NL; ++li, ++c) {
        mbar_wait(bar_d, (uint32_t)(c & 1));
        tc_fence_after();
        const uint32_t t_d = t_T[li & 1], t_a = t_T[(li + 1) & 1];  // A region of an odd layer = T[(li-1)&1]
        if (li == 0) {        // Dense_1 -> ConditionalNormalization_1 + gelu, + sample_effect   (reference _module.py:190-199)
          enc_ep<2>(t_d, 0u, half, row, xch, 1.0f, tail.bl[0], tail.gamma1 + (size_t)sid * 128, tail.beta1 + (size_t)sid * 128,
                    tail.sample_effect + (size_t)sid * 128);
        } else if (li & 1) {  // ResnetBlock Dense_0: relu(LN(.)) + h                            (reference _components.py:41-49)
          enc_ep<1>(t_d, t_a, half, row, xch, 1.0f, tail.bl[li], tail.lg[li], tail.lb[li], nullptr);
        } else {              // ResnetBlock Dense_out: relu(LN(.))                               (reference _components.py:49-51)
          enc_ep<1>(t_d, 0u, half, row, xch, 1.0f, tail.bl[li], tail.lg[li], tail.lb[li], nullptr);
        }
        tc_fence_before();
        mbar_arrive(bar_a);
      }
      // ---- mean head: u, LN(u), query tokens, z_base                                          (reference _components.py:95)
      mbar_wait(bar_d, (uint32_t)(c & 1));
      ++c;
      tc_fence_after();
      const int Lq = tail.Lq, L = tail.L;
      if (half == 0) {
        // u = D[:, 0:Lq] + b ; LN(u) ; query tokens = LN(u) . Wq            (reference _module.py:142, _components.py:195-197)
        float u[32];
        tmem_ld32(t_T[NL & 1], u);
        tmem_ld_wait();
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int l = 0; l < 32; ++l) {
          u[l] = (l < Lq) ? u[l] + __ldg(tail.bhead + l) : 0.f;
          s1 += u[l];
          s2 = fmaf(u[l], u[l], s2);
        }
        const float mean = s1 / (float)Lq;
        const float var = fmaxf(s2 / (float)Lq - mean * mean, 0.f);
        const float rstd = 1.0f / sqrtf(var + kLnEps);
        if (valid) {
          float qt[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) qt[i] = 0.f;
#pragma unroll
          for (int l = 0; l < 32; ++l)
            if (l < Lq) {
              const float ul = (u[l] - mean) * (rstd * __ldg(tail.uln_g + l)) + __ldg(tail.uln_b + l);
              cb.u[cell * Lq + l] = u[l];
              cb.u_ln[cell * Lq + l] = ul;
              const float4* wr = reinterpret_cast<const float4*>(tail.wq + l * 16);
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                const float4 w = __ldg(wr + i4);
                qt[4 * i4] = fmaf(ul, w.x, qt[4 * i4]); qt[4 * i4 + 1] = fmaf(ul, w.y, qt[4 * i4 + 1]);
                qt[4 * i4 + 2] = fmaf(ul, w.z, qt[4 * i4 + 2]); qt[4 * i4 + 3] = fmaf(ul, w.w, qt[4 * i4 + 3]);
              }
              if (!tail.has_zbase) cb.zbase[cell * L + l] = u[l];
            }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(cb.q_tok + cell * 16)[i] = make_float4(qt[4 * i], qt[4 * i + 1], qt[4 * i + 2], qt[4 * i + 3]);
        }
      } else if (tail.has_zbase) {
        // z_base = D[:, Lq:Lq+L] + b   (reference _module.py:166-168, folded into the head GEMM)
        for (int c0 = 0; c0 < tail.Nh; c0 += 32) {
          float v[32];
          tmem_ld32(t_T[NL & 1] + c0, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              const int j = c0 + q - Lq;
              if (j >= 0 && j < L) cb.zbase[cell * L + j] = v[q] + __ldg(tail.bhead + c0 + q);
            }
          }
        }
      }
      tc_fence_before();
      asm volatile("bar.sync 1, 256;" ::: "memory");  // head's TMEM reads are done before the next tile reuses T
    }
  } else if (warp_u == 16) {
    // ------------------------------------------------------------ phase-1 MMA issuer (warp-uniform branch, elect.sync
    // around the MMAs: 42-61 clk per tcgen05.mma instead of ~92 from `if (lane == 0)` code, see elect_one)
    {
      const uint32_t idesc = umma_idesc_f16(128);
      const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_t, 0);
      int64_t g = 0;
      for (int64_t ti = 0; ti < my_tiles; ++ti) {
        const uint32_t t_acc = tmem_base + (uint32_t)(ti & 1) * 128;
        if (ti >= 2) { mbar_wait(acc_free + (ti & 1), (uint32_t)(((ti >> 1) - 1) & 1)); tc_fence_after(); }
        for (int kb = 0; kb < nkb; ++kb, ++g) {
          const int s = (int)(g % (TMAX ? kEncRA : kEncPStages));
          const int sw = (int)(g % kEncRW);
          mbar_wait(full + s, (uint32_t)((g / (TMAX ? kEncRA : kEncPStages)) & 1));
          ENC_STAMP(g, 6);
          if (TMAX) mbar_wait(w_full + sw, (uint32_t)((g / kEncRW) & 1));
          tc_fence_after();
          ENC_STAMP(g, 7);
          const uint32_t a_h = TMAX ? smem_u32(aring + (size_t)s * 2 * kEncAHalf) : smem_u32(smem + s * kEncStageBytes);
          const uint32_t a_l = a_h + kEncAHalf;
          const uint32_t b_h = TMAX ? smem_u32(w0ring + (size_t)sw * kEncBBytes) : a_h + 2 * kEncAHalf, b_l = b_h + 4 * 2048;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t ah = umma_desc(a_h + ks * 2 * kEncALbo, kEncALbo, 128);
              const uint64_t al = umma_desc(a_l + ks * 2 * kEncALbo, kEncALbo, 128);
              const uint64_t bh = umma_desc(b_h + ks * 2 * 2048, 2048, 128);
              const uint64_t bl = umma_desc(b_l + ks * 2 * 2048, 2048, 128);
              umma_f16(t_acc, ah, bh, idesc, (kb | ks) ? 1u : 0u);
              umma_f16(t_acc, al, bh, idesc, 1u);
              umma_f16(t_acc, ah, bl, idesc, 1u);
            }
            umma_commit(empty + s);
            if (TMAX) umma_commit(w_empty + sw);
          }
          __syncwarp();
          ENC_STAMP(g, 8);
        }
        if (elect_one()) umma_commit(acc_full + (ti & 1));
        __syncwarp();
      }
    }
  } else if (warp_u == 17) {
    // ------------------------------------------------------------ tail MMA issuer: A from TMEM (warp-uniform, as above)
    {
      const uint32_t s_wb = smem_u32(wbuf);
      const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_t, 0);
      int64_t c = 0;
      for (int64_t ti = 0; ti < my_tiles; ++ti) {
        for (int li = 0; li <= NL; ++li, ++c) {
          const int N = li < NL ? 128 : tail.Nh;
          const uint32_t idn = umma_idesc_f16(N), lbo = (uint32_t)N * 16;
          const uint32_t w_hi = s_wb, w_lo = s_wb + (uint32_t)N * 128 * 2;
          mbar_wait(bar_w, (uint32_t)(c & 1));
          mbar_wait(bar_a, (uint32_t)(c & 1));
          tc_fence_after();
          const uint32_t t_a = li == 0 ? tmem_base + (uint32_t)(ti & 1) * 128 : tmem_base + 256 + (uint32_t)((li - 1) & 1) * 128;
          const uint32_t t_d = tmem_base + 256 + (uint32_t)(li & 1) * 128;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint32_t a = t_a + ks * 16;  // [8 hi words | 8 lo words] per 16-column chunk
              const uint64_t bh = umma_desc(w_hi + ks * 2 * lbo, lbo, 128), bl = umma_desc(w_lo + ks * 2 * lbo, lbo, 128);
              umma_f16_ts(t_d, a, bh, idn, ks == 0 ? 0u : 1u);
              umma_f16_ts(t_d, a + 8, bh, idn, 1u);
              umma_f16_ts(t_d, a, bl, idn, 1u);
            }
            umma_commit(bar_d);
            if (li == 0) umma_commit(acc_free + (ti & 1));  // the accumulator region has been consumed
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 18) {
    // ------------------------------------------------------------ W0 loader (TMA bulk copies)
    if (TMAX) {
      const int64_t total = my_tiles * nkb;
      if (lane == 0) {
        // raw X tiles: one tensor copy per K block
        for (int64_t g = 0; g < total; ++g) {
          const int64_t ti = g / nkb;
          const int kb = (int)(g - ti * nkb);
          const int64_t cell0 = ((int64_t)blockIdx.x + ti * gridDim.x) * 128;
          const int sx = (int)(g % kEncRX);
          ENC_STAMP(g, 9);
          if (g >= kEncRX) mbar_wait(x_empty + sx, (uint32_t)((g / kEncRX - 1) & 1));
          ENC_STAMP(g, 10);
          mbar_expect_tx(x_full + sx, kEncRawBytes);
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                  smem_u32(rawx + (size_t)sx * kEncRawBytes)),
              "l"(reinterpret_cast<uint64_t>(&xmap)), "r"(kb * kEncBK), "r"((int)cell0), "r"(smem_u32(x_full + sx))
              : "memory");
        }
      } else if (lane == 16) {
        for (int64_t g = 0; g < total; ++g) {
          const int sw = (int)(g % kEncRW);
          const int kb = (int)(g % nkb);
          if (g >= kEncRW) mbar_wait(w_empty + sw, (uint32_t)((g / kEncRW - 1) & 1));
          mbar_expect_tx(w_full + sw, kEncBBytes);
          bulk_g2s(w0ring + (size_t)sw * kEncBBytes, enc.w0img + (size_t)kb * kEncBBytes, kEncBBytes, w_full + sw);
        }
      }
    } else if (lane == 0) {
      const int64_t total = my_tiles * nkb;
      for (int64_t g = 0; g < total; ++g) {
        const int s = (int)(g % kEncPStages);
        const int64_t use = g / kEncPStages;
        const int kb = (int)(g % nkb);
        if (use > 0) mbar_wait(empty + s, (uint32_t)((use - 1) & 1));
        mbar_expect_tx(full + s, kEncBBytes);
        bulk_g2s(smem + s * kEncStageBytes + 2 * kEncAHalf, enc.w0img + (size_t)kb * kEncBBytes, kEncBBytes, full + s);
      }
    }
  } else {
    // ------------------------------------------------------------ tail weight loader (TMA bulk copies)
    if (lane == 0) {
      int64_t c = 0;
      for (int64_t ti = 0; ti < my_tiles; ++ti) {
        for (int li = 0; li <= NL; ++li, ++c) {
          if (c > 0) mbar_wait(bar_d, (uint32_t)((c - 1) & 1));  // previous layer no longer reads the buffer
          const uint32_t bytes = li < NL ? kEncLayerBytes : (uint32_t)(2 * tail.Nh * 128 * 2);
          const uint8_t* src = li < NL ? tail.wl[li] : tail.whead;
          mbar_expect_tx(bar_w, bytes);
          for (uint32_t o = 0; o < bytes; o += 8192) bulk_g2s(wbuf + o, src + o, 8192, bar_w);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem_base, 512);
}

// ---- host side: pack the tail weights
inline uint32_t enc_pack_hilo_d(std::vector<uint8_t>& img, const std::vector<double>& W /*[K=128][N]*/, int N, int Npad) {
  const uint32_t off = (uint32_t)img.size();
  const size_t half_elems = (size_t)Npad * 128;
  img.resize(off + 2 * half_elems * 2, 0);
  __half* hi = reinterpret_cast<__half*>(img.data() + off);
  __half* lo = hi + half_elems;
  for (int k = 0; k < 128; ++k)
    for (int n = 0; n < N; ++n) {
      const float w = (float)W[(size_t)k * N + n];
      const __half h = __float2half_rn(w);
      const size_t e = (size_t)(k / 8) * ((size_t)Npad * 8) + (size_t)n * 8 + (k % 8);
      hi[e] = h;
      lo[e] = __float2half_rn((float)(W[(size_t)k * N + n] - (double)__half2float(h)));
    }
  return off;
}

inline uint32_t enc_pack_hilo(std::vector<uint8_t>& img, const float* W /*[K=128][N]*/, int N, int Npad) {
  // canonical K-major [Npad x 128] f16, hi image then lo image
  const uint32_t off = (uint32_t)img.size();
  const size_t half_elems = (size_t)Npad * 128;
  img.resize(off + 2 * half_elems * 2, 0);
  __half* hi = reinterpret_cast<__half*>(img.data() + off);
  __half* lo = hi + half_elems;
  for (int k = 0; k < 128; ++k)
    for (int n = 0; n < N; ++n) {
      const float w = W[(size_t)k * N + n];
      const __half h = __float2half_rn(w);
      const size_t e = (size_t)(k / 8) * ((size_t)Npad * 8) + (size_t)n * 8 + (k % 8);
      hi[e] = h;
      lo[e] = __float2half_rn(w - __half2float(h));
    }
  return off;
}

inline int enc_tail_build(EncTail& t, const NetW& net, const tc_detail::PMap& pm, DeviceArena& arena) {
  using tc_detail::P;
  t.ready = false;
  if (net.H != 128 || net.Lq > 31) return 0;
  std::vector<uint8_t> img;
  std::vector<uint32_t> offs;
  offs.push_back(enc_pack_hilo(img, P(pm, "qu/Dense_1/kernel"), 128, 128));
  for (int b = 0; b < net.n_enc_blocks; ++b) {
    const std::string rb = "qu/NormalDistOutputNN_0/ResnetBlock_" + std::to_string(b);
    offs.push_back(enc_pack_hilo(img, P(pm, rb + "/Dense_0/kernel"), 128, 128));
    offs.push_back(enc_pack_hilo(img, P(pm, rb + "/Dense_1/kernel"), 128, 128));
  }
  // head: h -> [u | z_base]; z_base = u W_zb + b_zb = h (W_mean W_zb) + (b_mean W_zb + b_zb), folded in float64
  const int Lq = net.Lq, L = net.L;
  const int n_head = Lq + (net.has_zbase ? L : 0);
  const int Nh = ((n_head + 31) / 32) * 32;
  if (Nh > kEncHeadMaxN) return 0;
  std::vector<double> Wh((size_t)128 * n_head, 0.0);
  std::vector<float> bh(Nh, 0.f);
  {
    const float* Wm = P(pm, "qu/NormalDistOutputNN_0/Dense_0/kernel");  // [128][Lq]
    const float* bm = P(pm, "qu/NormalDistOutputNN_0/Dense_0/bias");
    for (int k = 0; k < 128; ++k)
      for (int l = 0; l < Lq; ++l) Wh[(size_t)k * n_head + l] = Wm[(size_t)k * Lq + l];
    for (int l = 0; l < Lq; ++l) bh[l] = bm[l];
    if (net.has_zbase) {
      const float* Wz = P(pm, "qz/Dense_0/kernel");  // [Lq][L]
      const float* bz = P(pm, "qz/Dense_0/bias");
      for (int k = 0; k < 128; ++k)
        for (int j = 0; j < L; ++j) {
          double a = 0.0;
          for (int l = 0; l < Lq; ++l) a += (double)Wm[(size_t)k * Lq + l] * Wz[(size_t)l * L + j];
          Wh[(size_t)k * n_head + Lq + j] = a;
        }
      for (int j = 0; j < L; ++j) {
        double a = bz[j];
        for (int l = 0; l < Lq; ++l) a += (double)bm[l] * Wz[(size_t)l * L + j];
        bh[Lq + j] = (float)a;
      }
    }
  }
  const uint32_t off_head = enc_pack_hilo_d(img, Wh, n_head, Nh);
  const uint8_t* dimg = nullptr;
  int err;
  if ((err = tc_upload(arena, img.data(), img.size(), (const void**)&dimg))) return err;
  t.n_layers = 1 + 2 * net.n_enc_blocks;
  for (int i = 0; i < t.n_layers; ++i) t.wl[i] = dimg + offs[i];
  t.whead = dimg + off_head;
  t.gamma0 = net.gamma[0]; t.beta0 = net.beta[0]; t.gamma1 = net.gamma[1]; t.beta1 = net.beta[1];
  t.sample_effect = net.sample_effect;
  t.bl[0] = net.enc1.b; t.lg[0] = nullptr; t.lb[0] = nullptr;
  for (int b = 0; b < net.n_enc_blocks; ++b) {
    const ResnetW& rb = net.enc_blocks[b];
    t.bl[1 + 2 * b] = rb.d0.b; t.lg[1 + 2 * b] = rb.ln0.scale; t.lb[1 + 2 * b] = rb.ln0.bias;
    t.bl[2 + 2 * b] = rb.dout.b; t.lg[2 + 2 * b] = rb.ln1.scale; t.lb[2 + 2 * b] = rb.ln1.bias;
  }
  t.Nh = Nh;
  t.has_zbase = net.has_zbase;
  if ((err = tc_upload(arena, bh.data(), bh.size() * sizeof(float), (const void**)&t.bhead))) return err;
  t.uln_g = net.u_ln.scale; t.uln_b = net.u_ln.bias;
  // query-token and z_base weights as dense [Lq][16] / [Lq][L] fp32 (the packed fp32 copies are padded)
  std::vector<float> wq((size_t)net.Lq * 16);
  const float* wqs = P(pm, "qz/AttentionBlock_0/DenseGeneral_0/kernel");
  for (size_t i = 0; i < wq.size(); ++i) wq[i] = wqs[i];
  if ((err = tc_upload(arena, wq.data(), wq.size() * sizeof(float), (const void**)&t.wq))) return err;
  t.Lq = net.Lq; t.L = net.L; t.S = net.S;
  { int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); t.sm_count = sms; }
  if (cudaFuncSetAttribute(k_encoder_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin) != cudaSuccess ||
      cudaFuncSetAttribute(k_encoder_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin) != cudaSuccess) {
    g_tc_error = "cudaFuncSetAttribute(k_encoder_tc) failed";
    return MRVI_ERR_CUDA;
  }
  t.ready = true;
  return 0;
}

// 2-D tensor map of X [n rows][G columns], row stride ldx floats, box = 32 columns x 128 rows (driver entry point
// fetched through the runtime: no link-time dependency on libcuda)
inline bool enc_make_xmap(const float* X, int64_t n, int G, int64_t ldx, CUtensorMap* out) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeFn>(p);
  }();
  if (!fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)G, (cuuint64_t)n};
  const cuuint64_t strides[1] = {(cuuint64_t)ldx * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kEncBK, 128u};
  const cuuint32_t estr[2] = {1u, 1u};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline void enc_full_tc_launch(const EncTc& enc, const EncTail& tail, const float* X, int64_t ldx, const int32_t* sid, int64_t n,
                               int G, const CellBuf& cb, cudaStream_t st) {
  const int64_t n_tiles = (n + 127) / 128;
  const int grid = (int)std::min<int64_t>(n_tiles, tail.sm_count);
  if (grid == 0) return;
  const int vec_ok = ((ldx % 4) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  // rows that are 16-byte aligned with a 16-byte multiple of columns stream through the copy engine
  static const bool no_tmax = getenv("MRVI_B200_ENCODER_LDG") != nullptr;
  CUtensorMap xmap;
  memset(&xmap, 0, sizeof(xmap));
  if (vec_ok && n < (1ll << 31) && !no_tmax && enc_make_xmap(X, n, G, ldx, &xmap))
    k_encoder_tc<true><<<grid, kEncPThreads, kEncTSmem, st>>>(X, ldx, sid, n, G, enc, tail, cb, vec_ok, xmap);
  else
    k_encoder_tc<false><<<grid, kEncPThreads, kEncPSmem, st>>>(X, ldx, sid, n, G, enc, tail, cb, vec_ok, xmap);
}

}  // namespace mrvi
