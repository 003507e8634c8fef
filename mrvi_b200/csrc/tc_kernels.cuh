// tcgen05 tensor-core path of the q(z|u,s) chain (MRVI_PRECISION_TC), sm_100a only.
//
// One CTA = 128 threads = one 128-row tile of (cell, sample) rows; thread r owns row r, which is
// TMEM lane r, so LayerNorm statistics are thread-local.  The chain of the reference
// (_components.py:181-230) is executed as five GEMMs on the 5th-gen tensor cores
// (tcgen05.mma, kind::f16, fp32 accumulators in TMEM) with CUDA-core epilogues in between:
//
//   attention  -> m[h,i] = softmax_j(t[h,i] kv_s[j]) . kv_s        (exact collapse of the 16x16x2 MHA;
//                 score is rank-1 in the scalar tokens, value/out projections fold into GEMM1)
//   GEMM1  D1[128x128] = A1[m | 1]            x B1      (MLP_0 Dense_0, attention out-proj folded in)
//   epi1   T = gelu(LN(D1))                              (fp32, thread-local row statistics)
//   GEMM2  D2[128x32]  = T x B2t + A1 x B2a              (Dense_2; the skip Dense_1 folded: (a Ws) Wout)
//   epi2   t2 = gelu(LN(D2))
//   GEMM3  D3[128x128] = A3[t2 | 1 | u_ln]    x B3      (MLP_0 Dense_0(out) and MLP_1 Dense_0 fused)
//   epi3   T = gelu(LN(D3))
//   GEMM4  D4[128x32]  = T x B4t + A3 x B4a
//   epi4   t4 = gelu(LN(D4))
//   GEMM5  D5[128xN5]  = A5[t4 | 1]           x B5      (MLP_1 final Dense, first L outputs)
//
// Every operand is an (hi, lo) pair of f16 values (hi = 11 leading bits, lo = the remainder) and every
// product is the 3-term split  Ah.Wh + Al.Wh + Ah.Wl  (fp32 accumulate), i.e. ~21 significant bits on
// both sides.  Round 1 rounded the weights to f16 once: a systematic 3-5e-4 normwise effect on the
// distances with a tail above the 1e-3 bar (6 of 120 random configurations); with the lo weight image
// the emulation (oracle/tc_emulation.py) and the measured error drop to the 1e-5 level.  All folding of
// weights is done in float64 on the host at create time (tc_build).
//
// Operands use the canonical no-swizzle K-major UMMA layout: 16-byte chunk kc of row r of an
// [R x K] operand lives at  base + kc * (R * 16) + r * 16  (LBO = R*16 bytes between K chunks,
// SBO = 128 bytes between groups of 8 rows), so thread r writes its row with conflict-free
// 16-byte shared stores.
#pragma once
#include <cuda_fp16.h>

#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace mrvi {

constexpr int kTcThreads = 128;
constexpr int kTcRows = 128;
constexpr float kLog2e = 1.4426950408889634f;

// epilogue vectors, passed by value so they are read as constant-bank operands
struct TcLn {
  float g1[128], b1[128];  // MLP_0 ResnetBlock LayerNorm_0
  float g2[32], b2[32];    // MLP_0 ResnetBlock LayerNorm_1
  float g3[128], b3[128];  // MLP_1 ResnetBlock LayerNorm_0
  float g4[32], b4[32];    // MLP_1 ResnetBlock LayerNorm_1
};

// One B operand of the chain inside the weight image: byte offsets of its hi / lo f16 copies and the byte stride
// between 16-byte K chunks (canonical K-major layout).  The 32-wide operands are stored CONCATENATED along N,
// [W_hi | W_lo] (64 rows per K chunk, lbo = 1024, lo = hi + 512): one N = 64 MMA then forms A.W_hi and A.W_lo side by
// side (tcgen05.mma costs ~42 clk per instruction for N <= 64, tools/mma_microbench.cu), while N = 32 descriptors on
// `h` / `l` still address the two copies separately.
struct TcOp {
  uint32_t h = 0, l = 0, lbo = 0;
};

struct TcNet {
  const float* kvtok = nullptr;   // [S][16] per-sample kv tokens (fp32)
  const float* kvl = nullptr;     // [S][16] kv * log2(e)
  const float* kvref = nullptr;   // [S][2]  (max_j kv, min_j kv) * log2(e)
  const uint8_t* wimg = nullptr;  // packed f16 operand image (device), copied verbatim to smem
  uint32_t wimg_bytes = 0;
  TcOp b1, b2t, b2a, b3, b4t, b4a, b5;  // operands inside the image
  int K1 = 48, K3 = 48, K5 = 48, N5 = 32;
  float alpha[2] = {0, 0}, beta[2] = {0, 0};  // t[h,i] = alpha[h] * q_tok[i] + beta[h]
  // Attention table (see tc_build): m_s(t) = sum_j softmax_j(t kv_s[j]) kv_s[j] is ONE smooth scalar function per sample;
  // att_tab[k][s] = coefficients (a0, a1, a2, a3) of the cubic Hermite interpolant of interval k (a(u) = a0 + u (a1 + u (a2 + u a3)),
  // u in [0, 1)) of a uniform grid over the whole reachable range of t.
  const float4* att_tab = nullptr;  // [att_n][S], nullptr = evaluate the softmax directly (16 ex2 per feature)
  int att_n = 0;                    // intervals
  float att_t0 = 0.f, att_scale = 0.f, att_xmax = 0.f;  // x = (t - t0) * scale clamped to [0, xmax], xmax < att_n
  float att_err = 0.f;              // verified max interpolation error / max_j |kv_s[j]|
  TcLn ln;
  int sm_count = 148;
  size_t smem_bytes = 0;
  size_t fused_smem_bytes = 0;  // 0: fused kernel unavailable for this shape
  bool ready = false;
};

static std::string g_tc_error;
inline const char* tc_error() { return g_tc_error.c_str(); }

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// SMEM matrix descriptor, K-major, no swizzle: LBO = byte distance between 16-byte K chunks,
// SBO = byte distance between 8-row groups.  (cute::UMMA::SmemDescriptor bit layout, version 1.)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

// instruction descriptor for kind::f16: A = B = f16, D = f32, both K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_f16(int N) {
  return (1u << 4)                     // c_format = F32
         | (0u << 7) | (0u << 10)      // a_format = b_format = F16
         | (0u << 15) | (0u << 16)     // a_major = b_major = K
         | ((uint32_t)(N >> 3) << 17)  // n_dim
         | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> 32 lanes x 32 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// A operand from TMEM (lane = row, 32-bit column = two consecutive K values), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// one lane of a converged warp (elect.sync): `if (warp_uniform_condition) { if (elect_one()) { tcgen05.mma ... } }` lets
// the compiler keep the MMA descriptors in uniform registers -- issued from `if (lane == 0)` code every UTCHMMA is
// wrapped in an ELECT / R2UR waterfall loop and costs ~92 clk instead of 42-61 (tools/mma_microbench.cu)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// gelu(tanh) with MUFU.TANH (measured max abs error 7.6e-6 on B200, tools/microbench.cu)
__device__ __forceinline__ float gelu_fast(float x) {
  const float c1 = 0.7978845608028654f, c2 = 0.7978845608028654f * 0.044715f;
  const float u = x * x;
  const float p = fmaf(u, c2, c1);
  const float t = tanh_approx(x * p);
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
// TWICE gelu(tanh) for the chain kernels: the 1/2 is folded into the next GEMM's weight rows (tc_build)
__device__ __forceinline__ float gelu2x_fast(float x) {
  const float c1 = 0.7978845608028654f, c2 = 0.7978845608028654f * 0.044715f;
  const float u = x * x;
  const float p = fmaf(u, c2, c1);
  const float t = tanh_approx(x * p);
  return fmaf(x, t, x);
}
// (hi, lo) f16 split of two floats, returned as packed f16x2 words:
// hi = RN_f16(x) (one packed conversion for the pair), lo = RN_f16(x - hi) with the remainder formed by the mixed-precision
// FMA (FHFMA: f32 = f16 * f16 + f32, the half selected by an operand modifier): 4 instructions per pair instead of 5-6
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  float la, lb;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));  // upper half = first source
  asm("{\n\t.reg .b16 l, u, m;\n\t"
      "mov.b32 {l, u}, %2;\n\t"
      "mov.b16 m, 0xBC00;\n\t"   // -1.0
      "fma.rn.f32.f16 %0, l, m, %3;\n\t"
      "fma.rn.f32.f16 %1, u, m, %4;\n\t}"
      : "=f"(la), "=f"(lb) : "r"(hi), "f"(a), "f"(b));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(lb), "f"(la));
}

// ------------------------------------------------------------------------------------------
// shared memory map (bytes, from a 1024-aligned base)
// ------------------------------------------------------------------------------------------
struct TcSmemMap {
  uint32_t w;        // weight image
  uint32_t a1h, a1l; // [128 x K1]   (also A5 = [t4 | 1])
  uint32_t a3h, a3l; // [128 x K3]
  uint32_t bars;     // 2 mbarriers + tmem slot
  uint32_t total;
};
__host__ __device__ inline TcSmemMap tc_smem_map(uint32_t wimg_bytes, int K1, int K3) {
  TcSmemMap m;
  uint32_t o = 0;
  auto take = [&](uint32_t bytes) { uint32_t r = o; o += (bytes + 1023u) & ~1023u; return r; };
  m.w = take(wimg_bytes);
  m.a1h = take(128 * K1 * 2); m.a1l = take(128 * K1 * 2);
  m.a3h = take(128 * K3 * 2); m.a3l = take(128 * K3 * 2);
  m.bars = take(64);
  m.total = o;
  return m;
}

// LayerNorm + gelu of NV = 32*NCH accumulator columns held in registers; writes (hi, lo) f16 chunks.
// dst_h / dst_l: this row's first 16-byte chunk; chunk stride = 128 rows * 16 B = 2048 B.
template <int NCH>
__device__ __forceinline__ void ln_gelu_store(float (&v)[NCH * 32], const float* __restrict__ g,
                                              const float* __restrict__ b, uint8_t* dst_h, uint8_t* dst_l) {
  constexpr int NV = NCH * 32;
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) { s1 += v[i]; s2 = fmaf(v[i], v[i], s2); }
  const float mean = s1 * (1.0f / NV);
  const float var = fmaxf(fmaf(s2, 1.0f / NV, -mean * mean), 0.f);
  const float rstd = rsqrt_approx(var + kLnEps);
  const float nb = -mean * rstd;
#pragma unroll
  for (int c = 0; c < NV / 8; ++c) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i0 = c * 8 + 2 * j, i1 = i0 + 1;
      const float y0 = gelu2x_fast(fmaf(fmaf(v[i0], rstd, nb), g[i0], b[i0]));
      const float y1 = gelu2x_fast(fmaf(fmaf(v[i1], rstd, nb), g[i1], b[i1]));
      split2(y0, y1, h[j], l[j]);
    }
    *reinterpret_cast<uint4*>(dst_h + c * 2048) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(dst_l + c * 2048) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// LayerNorm + gelu of the 128 accumulator columns of this row, written back IN PLACE to TMEM as the
// A operand of the next GEMM: f16 pairs, hi copy in columns [0, 64), lo copy in columns [64, 128).
// (The row's accumulators are all in registers before the first store, and a lane is private to
// its thread, so overwriting is safe.)
__device__ __forceinline__ void ln_gelu_store_tmem(float (&v)[128], const float* __restrict__ g,
                                                   const float* __restrict__ b, uint32_t t_lane) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 128; ++i) { s1 += v[i]; s2 = fmaf(v[i], v[i], s2); }
  const float mean = s1 * (1.0f / 128);
  const float var = fmaxf(fmaf(s2, 1.0f / 128, -mean * mean), 0.f);
  const float rstd = rsqrt_approx(var + kLnEps);
  const float nb = -mean * rstd;
#pragma unroll
  for (int c = 0; c < 2; ++c) {  // 64 values -> 32 hi words + 32 lo words
    uint32_t h[32], l[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const int i0 = c * 64 + 2 * j, i1 = i0 + 1;
      const float y0 = gelu2x_fast(fmaf(fmaf(v[i0], rstd, nb), g[i0], b[i0]));
      const float y1 = gelu2x_fast(fmaf(fmaf(v[i1], rstd, nb), g[i1], b[i1]));
      split2(y0, y1, h[j], l[j]);
    }
    tmem_st32(t_lane + c * 32, h);
    tmem_st32(t_lane + 64 + c * 32, l);
  }
  tmem_st_wait();
}

// [128 x 128] (A in TMEM: hi at a_tmem, lo at a_tmem + 64 columns) x [128 x N], weights (hi, lo): Ah.Wh + Al.Wh + Ah.Wl
__device__ __forceinline__ void issue_gemm_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t s_w, const TcOp& b, uint32_t idesc) {
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const uint64_t bd = umma_desc(s_w + b.h + ks * 2 * b.lbo, b.lbo, 128);
    const uint64_t bl = umma_desc(s_w + b.l + ks * 2 * b.lbo, b.lbo, 128);
    umma_f16_ts(d_tmem, a_tmem + ks * 8, bd, idesc, ks == 0 ? 0u : 1u);
    umma_f16_ts(d_tmem, a_tmem + 64 + ks * 8, bd, idesc, 1u);
    umma_f16_ts(d_tmem, a_tmem + ks * 8, bl, idesc, 1u);
  }
}

// issue the MMAs of one [128 x K] x [K x N] product: (hi, lo) activations against (hi, lo) weights, 3 terms per K slice.
// K_lo: K extent of the lo copy of A (A1 / A5 = [features | 1 | 0...]: the lo halves of the constant column and of the
// padding are zero, so their last K slice is skipped and not even stored).
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_h, uint32_t a_l, uint32_t s_w, const TcOp& b, int K, int K_lo,
                                           uint32_t idesc, bool first_clears) {
  const uint32_t a_lbo = 128 * 16;
  for (int ks = 0; ks < K / 16; ++ks) {
    const uint64_t bd = umma_desc(s_w + b.h + ks * 2 * b.lbo, b.lbo, 128);
    const uint64_t bl = umma_desc(s_w + b.l + ks * 2 * b.lbo, b.lbo, 128);
    const uint64_t ah = umma_desc(a_h + ks * 2 * a_lbo, a_lbo, 128);
    umma_f16(d_tmem, ah, bd, idesc, (first_clears && ks == 0) ? 0u : 1u);
    if (ks * 16 < K_lo) umma_f16(d_tmem, umma_desc(a_l + ks * 2 * a_lbo, a_lbo, 128), bd, idesc, 1u);
    umma_f16(d_tmem, ah, bl, idesc, 1u);
  }
}

// ------------------------------------------------------------------------------------------
// the chain kernel: persistent over 128-row tiles
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 2)
k_chain_tc(const __grid_constant__ TcNet tc, CellBuf cb, int64_t n_cells, int S, int L, int Lq,
           float* __restrict__ z_out, int64_t n_tiles, int cells_per_tile, int tiles_per_cell) {
  // no-swizzle UMMA operands and bulk copies need 16-byte alignment only; keeping the plain array (no integer
  // re-alignment) lets the compiler see the shared address space (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem_tc[];
  uint8_t* const smem = smem_tc;
  const TcSmemMap sm = tc_smem_map(tc.wimg_bytes, tc.K1, tc.K3);
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + sm.bars);
  uint64_t* bar_mma = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(bar_w, tc.wimg_bytes);
    bulk_g2s(smem + sm.w, tc.wimg, tc.wimg_bytes, bar_w);
  }
  const uint32_t s_w = smem_u32(smem + sm.w);
  const uint32_t s_a1h = smem_u32(smem + sm.a1h), s_a1l = smem_u32(smem + sm.a1l);
  const uint32_t s_a3h = smem_u32(smem + sm.a3h), s_a3l = smem_u32(smem + sm.a3l);
  uint8_t* row_a1h = smem + sm.a1h + tid * 16;
  uint8_t* row_a1l = smem + sm.a1l + tid * 16;
  uint8_t* row_a3h = smem + sm.a3h + tid * 16;
  uint8_t* row_a3l = smem + sm.a3l + tid * 16;
  const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);  // this warp's TMEM lane quarter
  const uint32_t idesc128 = umma_idesc_f16(128), idesc32 = umma_idesc_f16(32), idescN5 = umma_idesc_f16(tc.N5);
  uint32_t phase = 0;
  mbar_wait(bar_w, 0);  // weights landed (async proxy wrote them; MMA reads them through the same proxy)

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // ---- row -> (cell, s)
    int64_t cell;
    int s;
    bool valid;
    if (tiles_per_cell == 1) {
      const int lc = tid / S;
      cell = tile * cells_per_tile + lc;
      s = tid - lc * S;
      valid = lc < cells_per_tile && cell < n_cells;
    } else {
      cell = tile / tiles_per_cell;
      s = (int)(tile % tiles_per_cell) * kTcRows + tid;
      valid = s < S;
    }
    // the sampled paths re-map rows: per-row features and / or the cell's own sample tables (see CellBuf)
    const int64_t cc = valid ? (cb.per_row ? cell * S + s : cell) : 0;
    const int ss = valid ? (cb.cell_sample ? __ldg(cb.cell_sample + cell) : s) : 0;

    // ---- attention features m[h*16+i] -> A1 = [m | 1 | 0...], A3 static part = [.. | u_ln | 1 | 0..]
    {
      float kvl[16];
      const float4* kp = reinterpret_cast<const float4*>(tc.kvl + (size_t)ss * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t4 = __ldg(kp + q);
        kvl[4 * q] = t4.x; kvl[4 * q + 1] = t4.y; kvl[4 * q + 2] = t4.z; kvl[4 * q + 3] = t4.w;
      }
      const float rmax = __ldg(tc.kvref + ss * 2), rmin = __ldg(tc.kvref + ss * 2 + 1);
      float qt[16];
      const float4* qp = reinterpret_cast<const float4*>(cb.q_tok + cc * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 t4 = __ldg(qp + q);
        qt[4 * q] = t4.x; qt[4 * q + 1] = t4.y; qt[4 * q + 2] = t4.z; qt[4 * q + 3] = t4.w;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {  // 8 features per 16-byte chunk
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int j2 = 0; j2 < 4; ++j2) {
            float mm[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int i = c * 8 + j2 * 2 + e;
              const float t = fmaf(tc.alpha[h], qt[i], tc.beta[h]);
              const float off = -t * (t >= 0.f ? rmax : rmin);  // exponent <= 0 for every j
              float den = 0.f, num = 0.f;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float ex = ex2_approx(fmaf(t, kvl[j], off));
                den += ex;
                num = fmaf(ex, kvl[j], num);
              }
              mm[e] = num * rcp_approx(den);  // = log2(e) * m[h,i]; the 1/log2(e) is folded into B1 / B2a
            }
            split2(mm[0], mm[1], hw[j2], lw[j2]);
          }
          *reinterpret_cast<uint4*>(row_a1h + (h * 2 + c) * 2048) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          *reinterpret_cast<uint4*>(row_a1l + (h * 2 + c) * 2048) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
      // constant-one column (k = 32) and zero padding of A1
      const uint32_t one_h = 0x00003C00u;  // f16 1.0 in the low half
      for (int c = 4; c < tc.K1 / 8; ++c) {
        *reinterpret_cast<uint4*>(row_a1h + c * 2048) = make_uint4(c == 4 ? one_h : 0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(row_a1l + c * 2048) = make_uint4(0u, 0u, 0u, 0u);
      }
      // A3 static part: column 32 = 1.0, columns 33.. = u_ln[0..Lq), then zeros
      const float* ul = cb.u_ln + cc * Lq;
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
    __syncthreads();

    // ---- GEMM1
    if (tid == 0) {
      tc_fence_after();
      issue_gemm(tmem_base, s_a1h, s_a1l, s_w, tc.b1, tc.K1, tc.K1, idesc128, true);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, phase); phase ^= 1;
    tc_fence_after();
    {
      float v[128];
      tmem_ld32(t_lane + 0, v); tmem_ld32(t_lane + 32, v + 32); tmem_ld32(t_lane + 64, v + 64); tmem_ld32(t_lane + 96, v + 96);
      tmem_ld_wait();
      ln_gelu_store_tmem(v, tc.ln.g1, tc.ln.b1, t_lane);
    }
    tc_fence_before();
    __syncthreads();

    // ---- GEMM2: D2 = T x B2t + A1 x B2a
    if (tid == 0) {
      tc_fence_after();
      issue_gemm_ts(tmem_base + 128, tmem_base, s_w, tc.b2t, idesc32);
      issue_gemm(tmem_base + 128, s_a1h, s_a1l, s_w, tc.b2a, tc.K1, tc.K1, idesc32, false);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, phase); phase ^= 1;
    tc_fence_after();
    {
      float v[32];
      tmem_ld32(t_lane + 128, v);
      tmem_ld_wait();
      ln_gelu_store<1>(v, tc.ln.g2, tc.ln.b2, row_a3h, row_a3l);
    }
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();

    // ---- GEMM3
    if (tid == 0) {
      tc_fence_after();
      issue_gemm(tmem_base, s_a3h, s_a3l, s_w, tc.b3, tc.K3, tc.K3, idesc128, true);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, phase); phase ^= 1;
    tc_fence_after();
    {
      float v[128];
      tmem_ld32(t_lane + 0, v); tmem_ld32(t_lane + 32, v + 32); tmem_ld32(t_lane + 64, v + 64); tmem_ld32(t_lane + 96, v + 96);
      tmem_ld_wait();
      ln_gelu_store_tmem(v, tc.ln.g3, tc.ln.b3, t_lane);
    }
    tc_fence_before();
    __syncthreads();

    // ---- GEMM4: D4 = T x B4t + A3 x B4a
    if (tid == 0) {
      tc_fence_after();
      issue_gemm_ts(tmem_base + 128, tmem_base, s_w, tc.b4t, idesc32);
      issue_gemm(tmem_base + 128, s_a3h, s_a3l, s_w, tc.b4a, tc.K3, tc.K3, idesc32, false);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, phase); phase ^= 1;
    tc_fence_after();
    {
      float v[32];
      tmem_ld32(t_lane + 128, v);
      tmem_ld_wait();
      // A5 = [t4 | 1 | 0] reuses the A1 buffers (GEMM2 was their last reader); the one / zero columns are still in place
      ln_gelu_store<1>(v, tc.ln.g4, tc.ln.b4, row_a1h, row_a1l);
    }
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();

    // ---- GEMM5
    if (tid == 0) {
      tc_fence_after();
      issue_gemm(tmem_base, s_a1h, s_a1l, s_w, tc.b5, tc.K5, tc.K5, idescN5, true);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, phase); phase ^= 1;
    tc_fence_after();
    {
      float v[64];
      tmem_ld32(t_lane, v);
      if (tc.N5 > 32) tmem_ld32(t_lane + 32, v + 32);
      tmem_ld_wait();
      if (valid && z_out != nullptr) {
        const float* zb = cb.zbase + cc * L;
        float* dst = z_out + (cell * S + s) * (int64_t)L;
#pragma unroll
        for (int l = 0; l < 64; ++l)
          if (l < L) dst[l] = __ldg(zb + l) + v[l];
      }
    }
    tc_fence_before();
    __syncthreads();  // every thread has read D5 before the next tile's GEMM1 overwrites TMEM
  }

  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------
// host side: weight folding (float64) and packing into the f16 operand image
// ------------------------------------------------------------------------------------------
struct DeviceArena;

namespace tc_detail {
using PMap = std::map<std::string, std::pair<const float*, int64_t>>;
using Mat = std::vector<double>;  // row-major

inline const float* P(const PMap& m, const std::string& k) { return m.at(k).first; }

// C[K x N] = A[K x J] * B[J x N]
inline Mat matmul(const Mat& A, const Mat& B, int K, int J, int N) {
  Mat C((size_t)K * N, 0.0);
  for (int k = 0; k < K; ++k)
    for (int j = 0; j < J; ++j) {
      const double a = A[(size_t)k * J + j];
      for (int n = 0; n < N; ++n) C[(size_t)k * N + n] += a * B[(size_t)j * N + n];
    }
  return C;
}
inline Mat load(const float* p, size_t n) { return Mat(p, p + n); }

// append the [N x Kpad] K-major canonical f16 copies of B[K x N] (rows >= K are zero): hi = B rounded to f16, lo = the
// remainder B - hi rounded to f16 (subnormal for the smallest weights: absolute resolution 2^-24).  cat = false: two
// separate [Npad x Kpad] operands; cat = true: one [2 Npad x Kpad] operand [hi | lo] concatenated along N.
inline TcOp append_operand(std::vector<uint8_t>& img, const Mat& B, int K, int N, int Npad, int Kpad, bool cat) {
  TcOp op;
  const size_t one = (size_t)Npad * Kpad * 2;
  op.h = (uint32_t)img.size();
  img.resize(img.size() + 2 * one, 0);
  const int rows = cat ? 2 * Npad : Npad;            // operand rows per K chunk
  op.lbo = (uint32_t)rows * 16;
  op.l = cat ? op.h + (uint32_t)Npad * 16 : op.h + (uint32_t)one;
  __half* base = reinterpret_cast<__half*>(img.data() + op.h);
  const size_t lo_elems = (op.l - op.h) / 2;
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) {
      const size_t e = (size_t)(k / 8) * ((size_t)rows * 8) + (size_t)n * 8 + (k % 8);
      const double w = B[(size_t)k * N + n];
      const __half h = __float2half_rn((float)w);
      base[e] = h;
      base[lo_elems + e] = __float2half_rn((float)(w - (double)__half2float(h)));
    }
  while (img.size() % 1024) img.push_back(0);
  return op;
}
}  // namespace tc_detail

int tc_upload(DeviceArena& arena, const void* host, size_t bytes, const void** dev);  // defined in mrvi_b200.cu

inline int tc_build(TcNet& tc, const NetW& net, const MrviDims& d, const tc_detail::PMap& pm, DeviceArena& arena, int sm_count) {
  using namespace tc_detail;
  const int E = d.n_latent_sample, C = d.n_channels, nh = d.n_heads, M = d.qz_n_hidden, L = d.n_latent;
  const int Lq = d.n_latent_u > 0 ? d.n_latent_u : L;
  if (E != 16 || C != 4 || nh != 2 || M != 32 || d.qz_n_layers != 1 || Lq > 31 || L > 64) {
    g_tc_error = "MRVI_PRECISION_TC supports n_latent_sample=16, n_channels=4, n_heads=2, qz n_hidden=32, qz n_layers=1, "
                 "n_latent_u<=31, n_latent<=64 (use MRVI_PRECISION_FP32 for other hyper-parameters)";
    return MRVI_ERR_UNSUPPORTED;
  }
  const int R = kResnetHidden, hd = C, S = d.n_sample;
  const std::string ab = "qz/AttentionBlock_0", mha = ab + "/MultiHeadDotProductAttention_0";
  const float *wq = P(pm, mha + "/query/kernel"), *bq = P(pm, mha + "/query/bias");
  const float *wk = P(pm, mha + "/key/kernel"), *wv = P(pm, mha + "/value/kernel"), *bv = P(pm, mha + "/value/bias");
  const float *wo = P(pm, mha + "/out/kernel"), *bo = P(pm, mha + "/out/bias");
  const double inv = 1.0 / std::sqrt((double)hd), ln2 = std::log(2.0);
  for (int h = 0; h < nh; ++h) {
    double a = 0, b = 0;
    for (int dd = 0; dd < hd; ++dd) { a += (double)wq[h * hd + dd] * wk[h * hd + dd]; b += (double)bq[h * hd + dd] * wk[h * hd + dd]; }
    tc.alpha[h] = (float)(a * inv);
    tc.beta[h] = (float)(b * inv);
  }
  // o[i,c] = sum_h m[h,i] pv[h,c] + ro[c]
  Mat pv((size_t)nh * C, 0.0), ro(C, 0.0);
  for (int c = 0; c < C; ++c) {
    ro[c] = bo[c];
    for (int h = 0; h < nh; ++h)
      for (int dd = 0; dd < hd; ++dd) {
        pv[h * C + c] += (double)wv[h * hd + dd] * wo[(h * hd + dd) * C + c];
        ro[c] += (double)bv[h * hd + dd] * wo[(h * hd + dd) * C + c];
      }
  }
  const std::string b0 = ab + "/MLP_0/ResnetBlock_0", b1 = ab + "/MLP_1/ResnetBlock_0";
  const int K1r = nh * E + 1;  // [m (scaled by log2e in the kernel) | 1]
  auto fold_first = [&](const std::string& dense) {  // [K1r x R]: rows (h,i) then the bias row
    const float *W = P(pm, dense + "/kernel"), *bb = P(pm, dense + "/bias");
    Mat B((size_t)K1r * R, 0.0);
    for (int n = 0; n < R; ++n) {
      double bias = bb[n];
      for (int i = 0; i < E; ++i)
        for (int c = 0; c < C; ++c) {
          const double w = W[(size_t)(i * C + c) * R + n];
          bias += ro[c] * w;
          for (int h = 0; h < nh; ++h) B[(size_t)(h * E + i) * R + n] += pv[h * C + c] * w * ln2;  // kernel feeds log2e * m
        }
      B[(size_t)(K1r - 1) * R + n] = bias;
    }
    return B;
  };
  Mat B1 = fold_first(b0 + "/Dense_0");
  Mat Bs = fold_first(b0 + "/Dense_1");
  Mat Wout0 = load(P(pm, b0 + "/Dense_2/kernel"), (size_t)R * M);
  Mat B2a = matmul(Bs, Wout0, K1r, R, M);
  { const float* bb = P(pm, b0 + "/Dense_2/bias"); for (int n = 0; n < M; ++n) B2a[(size_t)(K1r - 1) * M + n] += bb[n]; }
  // A3 = [t2 (M) | 1 | u_ln (Lq)]   (the constant column sits at k = M in A1, A3 and A5 alike)
  const int K3r = M + Lq + 1;
  Mat Wfin0 = load(P(pm, ab + "/MLP_0/Dense_0/kernel"), (size_t)M * E);
  const float* bfin0 = P(pm, ab + "/MLP_0/Dense_0/bias");
  auto fold_second = [&](const std::string& dense) {  // [K3r x R]
    const float *W = P(pm, dense + "/kernel"), *bb = P(pm, dense + "/bias");  // (Lq+E, R)
    Mat We = load(W + (size_t)Lq * R, (size_t)E * R);
    Mat top = matmul(Wfin0, We, M, E, R);
    Mat B((size_t)K3r * R, 0.0);
    std::copy(top.begin(), top.end(), B.begin());
    for (int n = 0; n < R; ++n) {  // row M: the constant-one column (same position as in A1 / A5)
      double bias = bb[n];
      for (int e = 0; e < E; ++e) bias += (double)bfin0[e] * We[(size_t)e * R + n];
      B[(size_t)M * R + n] = bias;
    }
    for (int k = 0; k < Lq; ++k)
      for (int n = 0; n < R; ++n) B[(size_t)(M + 1 + k) * R + n] = W[(size_t)k * R + n];
    return B;
  };
  Mat B3 = fold_second(b1 + "/Dense_0");
  Mat S3 = fold_second(b1 + "/Dense_1");
  Mat Wout1 = load(P(pm, b1 + "/Dense_2/kernel"), (size_t)R * M);
  Mat B4a = matmul(S3, Wout1, K3r, R, M);
  { const float* bb = P(pm, b1 + "/Dense_2/bias"); for (int n = 0; n < M; ++n) B4a[(size_t)M * M + n] += bb[n]; }
  // A5 = [t4 | 1]; only the first L output columns are kept (use_map=False keeps the mean half)
  const int out_dim = d.use_map ? L : 2 * L;
  const float *Wf1 = P(pm, ab + "/MLP_1/Dense_0/kernel"), *bf1 = P(pm, ab + "/MLP_1/Dense_0/bias");
  const int K5r = M + 1;
  Mat B5((size_t)K5r * L, 0.0);
  for (int k = 0; k < M; ++k)
    for (int n = 0; n < L; ++n) B5[(size_t)k * L + n] = Wf1[(size_t)k * out_dim + n];
  for (int n = 0; n < L; ++n) B5[(size_t)M * L + n] = bf1[n];

  tc.K1 = 48; tc.K5 = 48;
  tc.K3 = round_up(K3r, 16);
  tc.N5 = round_up(L, 16) <= 32 ? 32 : 64;
  // The epilogues emit TWICE gelu (gelu_fast2 / gelu2x_fast): the factor 1/2 goes into the weight rows that multiply a
  // gelu output -- all of B2t / B4t (the 128-wide hidden activations) and the first M rows of B3, B4a, B5 (t2 / t4).
  // Exact: a power of two.
  {
    auto half_rows = [](Mat& B, int rows, int N) { for (size_t i = 0; i < (size_t)rows * N; ++i) B[i] *= 0.5; };
    half_rows(Wout0, R, M);
    half_rows(Wout1, R, M);
    half_rows(B3, M, R);
    half_rows(B4a, M, M);
    half_rows(B5, M, L);
  }
  std::vector<uint8_t> img;
  tc.b1 = append_operand(img, B1, K1r, R, R, tc.K1, false);
  tc.b2t = append_operand(img, Wout0, R, M, M, R, true);
  tc.b2a = append_operand(img, B2a, K1r, M, M, tc.K1, true);
  tc.b3 = append_operand(img, B3, K3r, R, R, tc.K3, false);
  tc.b4t = append_operand(img, Wout1, R, M, M, R, true);
  tc.b4a = append_operand(img, B4a, K3r, M, M, tc.K3, true);
  tc.b5 = append_operand(img, B5, K5r, L, tc.N5, tc.K5, false);
  tc.wimg_bytes = (uint32_t)img.size();
  int err;
  if ((err = tc_upload(arena, img.data(), img.size(), (const void**)&tc.wimg))) return err;

  auto cp = [&](float* dst, const std::string& name, int n) { const float* s = P(pm, name); for (int i = 0; i < n; ++i) dst[i] = s[i]; };
  cp(tc.ln.g1, b0 + "/LayerNorm_0/scale", R); cp(tc.ln.b1, b0 + "/LayerNorm_0/bias", R);
  cp(tc.ln.g2, b0 + "/LayerNorm_1/scale", M); cp(tc.ln.b2, b0 + "/LayerNorm_1/bias", M);
  cp(tc.ln.g3, b1 + "/LayerNorm_0/scale", R); cp(tc.ln.b3, b1 + "/LayerNorm_0/bias", R);
  cp(tc.ln.g4, b1 + "/LayerNorm_1/scale", M); cp(tc.ln.b4, b1 + "/LayerNorm_1/bias", M);

  // per-sample tables from the fp32 kv tokens (already on the device, computed by k_sample_tables)
  std::vector<float> kv((size_t)S * E);
  if (cudaMemcpy(kv.data(), tc.kvtok, kv.size() * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
    g_tc_error = "cudaMemcpy of kv tokens failed";
    return MRVI_ERR_CUDA;
  }
  std::vector<float> kvl((size_t)S * E), kvref((size_t)S * 2);
  for (int s = 0; s < S; ++s) {
    float mx = -INFINITY, mn = INFINITY;
    for (int j = 0; j < E; ++j) {
      const float v = kv[(size_t)s * E + j] * kLog2e;
      kvl[(size_t)s * E + j] = v;
      mx = std::max(mx, v); mn = std::min(mn, v);
    }
    kvref[s * 2] = mx; kvref[s * 2 + 1] = mn;
  }
  if ((err = tc_upload(arena, kvl.data(), kvl.size() * sizeof(float), (const void**)&tc.kvl))) return err;
  if ((err = tc_upload(arena, kvref.data(), kvref.size() * sizeof(float), (const void**)&tc.kvref))) return err;

  // ---- attention table.  The 16 x 16 x 2-head attention only enters the chain through the 32 softmax-weighted means
  // m_s(t[h,i]) = sum_j softmax_j(t kv_s[j]) kv_s[j]: one smooth, monotone scalar function per SAMPLE (the derivative of
  // log-sum-exp), evaluated at 32 points per (cell, sample) row -- 512 MUFU.EX2 per row, 27 % of the fused kernel's time.
  // It is tabulated here in float64 as a cubic-Hermite table on a uniform grid over the WHOLE reachable range of t:
  // q_i = LN(u).(g o Wq[:, i]) + b.Wq[:, i] with |LN(u)|_2 <= sqrt(Lq) and sum LN(u) = 0, so
  // |q_i - b.Wq[:, i]| <= sqrt(Lq) |c_i - mean(c_i)|_2 (c_i = g o Wq[:, i]) -- no input can leave the table.  The grid is
  // refined until the measured interpolation error (3 interior points of every interval of every sample) is below
  // 2^-22 of max_j |kv_s[j]|, i.e. below the (hi, lo) f16 resolution the features are fed to the tensor cores with; when
  // 4096 intervals do not reach it the table is dropped and the kernels evaluate the softmax directly.
  // Layout [interval][sample]: the rows of a cell share the cell's 32 values of t, so a warp reads 32 consecutive
  // 16-byte entries (coalesced, L2 resident: 16 B x intervals x S).
  if (!getenv("MRVI_B200_NO_ATT_TABLE")) {
    const float *g = P(pm, "qz/u_ln/scale"), *bb = P(pm, "qz/u_ln/bias"), *Wq = P(pm, ab + "/DenseGeneral_0/kernel");  // (Lq, E, 1)
    double T0 = INFINITY, T1 = -INFINITY;
    for (int i = 0; i < E; ++i) {
      double mean = 0, ctr = 0, ss = 0;
      for (int l = 0; l < Lq; ++l) { mean += (double)g[l] * Wq[(size_t)l * E + i]; ctr += (double)bb[l] * Wq[(size_t)l * E + i]; }
      mean /= Lq;
      for (int l = 0; l < Lq; ++l) { const double c = (double)g[l] * Wq[(size_t)l * E + i] - mean; ss += c * c; }
      const double rad = std::sqrt((double)Lq) * std::sqrt(ss);
      for (int h = 0; h < nh; ++h)
        for (double q : {ctr - rad, ctr + rad}) {
          const double t = (double)tc.alpha[h] * q + (double)tc.beta[h];
          T0 = std::min(T0, t); T1 = std::max(T1, t);
        }
    }
    const double pad = 1e-3 * (T1 - T0) + 1e-4;   // float32 rounding of q / t
    T0 -= pad; T1 += pad;
    auto feval = [&](int s, double t, double* deriv) {   // value (and derivative) in the kernel's units: kv * log2(e)
      double mx = -INFINITY;
      for (int j = 0; j < E; ++j) mx = std::max(mx, t * (double)kvl[(size_t)s * E + j]);
      double den = 0, num = 0, num2 = 0;
      for (int j = 0; j < E; ++j) {
        const double k = kvl[(size_t)s * E + j], w = std::exp2(t * k - mx);
        den += w; num += w * k; num2 += w * k * k;
      }
      const double m = num / den;
      if (deriv) *deriv = ln2 * (num2 / den - m * m);
      return m;
    };
    for (int N = 128; N <= 4096 && std::isfinite(T0) && T1 > T0; N *= 2) {
      const double hstep = (T1 - T0) / N;
      std::vector<float4> tab((size_t)N * S);
      double worst = 0;
      for (int s = 0; s < S && worst <= 2.4e-7; ++s) {
        double scale_s = 0;
        for (int j = 0; j < E; ++j) scale_s = std::max(scale_s, std::fabs((double)kvl[(size_t)s * E + j]));
        double d0, f0 = feval(s, T0, &d0);
        for (int k = 0; k < N; ++k) {
          double d1;
          const double f1 = feval(s, T0 + (k + 1) * hstep, &d1);
          // the cubic Hermite interpolant of the interval in monomial form (a0, a1, a2, a3), a(u) = a0 + u (a1 + u (a2 + u a3)):
          // three FMAs per look-up in the kernel instead of eight instructions
          const double hd0 = d0 * hstep, hd1 = d1 * hstep, dfd = f1 - f0;
          const float4 e = make_float4((float)f0, (float)hd0, (float)(3 * dfd - 2 * hd0 - hd1), (float)(hd0 + hd1 - 2 * dfd));
          tab[(size_t)k * S + s] = e;
          for (double u : {0.25, 0.5, 0.75, 1.0}) {   // the float32 entry, evaluated as the kernel does (u = 1: continuity)
            const double c2 = e.z, c3 = e.w;
            const double val = e.x + u * (e.y + u * (c2 + u * c3));
            worst = std::max(worst, std::fabs(val - feval(s, T0 + (k + u) * hstep, nullptr)) / std::max(scale_s, 1e-30));
          }
          f0 = f1; d0 = d1;
        }
      }
      if (worst <= 2.4e-7) {
        if ((err = tc_upload(arena, tab.data(), tab.size() * sizeof(float4), (const void**)&tc.att_tab))) return err;
        tc.att_n = N;
        tc.att_t0 = (float)T0;
        tc.att_scale = (float)(1.0 / hstep);
        tc.att_xmax = std::nextafter((float)N, 0.0f);
        tc.att_err = (float)worst;
        break;
      }
    }
  }

  tc.sm_count = sm_count;
  tc.smem_bytes = tc_smem_map(tc.wimg_bytes, tc.K1, tc.K3).total + 1024;
  if (tc.smem_bytes > 227 * 1024) {
    g_tc_error = "tensor-core chain needs more than 227 KiB of shared memory";
    return MRVI_ERR_UNSUPPORTED;
  }
  if (cudaFuncSetAttribute(k_chain_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin) != cudaSuccess) {
    g_tc_error = "cudaFuncSetAttribute(k_chain_tc) failed";
    return MRVI_ERR_CUDA;
  }
  tc.ready = true;
  return 0;
}

// launches the chain for n cells; per-cell inputs (q_tok, u_ln, zbase) must already be in cb
// rows_per_cell: 0 = net.S (mean path); the baseline of the sampled paths runs 2 * mc rows per cell
inline int tc_launch_qz(const TcNet& tc, const NetW& net, int64_t n, const CellBuf& cb, float* z, cudaStream_t st,
                        int rows_per_cell = 0) {
  const int S = rows_per_cell > 0 ? rows_per_cell : net.S;
  int cpt = 1, tpc = 1;
  int64_t n_tiles;
  if (S <= kTcRows) { cpt = kTcRows / S; n_tiles = (n + cpt - 1) / cpt; }
  else { tpc = (S + kTcRows - 1) / kTcRows; n_tiles = n * tpc; }
  const int grid = (int)std::min<int64_t>(n_tiles, 2 * (int64_t)tc.sm_count);
  if (grid == 0) return 0;
  k_chain_tc<<<grid, kTcThreads, tc.smem_bytes, st>>>(tc, cb, n, S, net.L, net.Lq, z, n_tiles, cpt, tpc);
  return 0;
}


// ==========================================================================================
// Encoder first layer on the tensor cores:  H1[n x H] = log1p(X)[n x G] . W0[G x H] + b0   (H = 128)
// (reference _module.py:189-190, the only large-K contraction of the path; HBM-bound on reading X).
//
// Warp-specialised, 3-stage mbarrier pipeline, one CTA = 128 cells:
//   warps 0-3  producers: coalesced LDG.128 of X, log1p (MUFU.LG2), (hi, lo) f16 split, conflict-free
//              16-byte shared stores in the K-major UMMA layout; afterwards the epilogue (TMEM -> H1)
//   warp 4     one elected thread issues tcgen05.mma:  per 16-wide K slice  Ah.Wh + Al.Wh + Ah.Wl
//              (f16 operands, fp32 accumulate: ~21 significant bits, see oracle/tc_emulation.py)
//   warp 5     one elected thread streams the pre-packed W0 image (hi|lo, 16 KiB per 32-wide K block)
//              with cp.async.bulk (TMA) into the stage's B buffer
// W0 is scaled by 2^8 before the split so that the lo halves stay in the f16 normal range; the
// epilogue multiplies by 2^-8 (exact).
// ==========================================================================================
constexpr int kEncBK = 32;
constexpr int kEncStages = 3;
constexpr int kEncALbo = 129 * 16;                       // 16-byte row pad per K-chunk plane: conflict-free stores
constexpr int kEncAHalf = 4 * kEncALbo;                  // one of (hi, lo): 4 K-chunks
constexpr int kEncBBytes = 2 * 4 * 128 * 16;             // hi | lo, 4 K-chunks x 128 n x 16 B
constexpr int kEncStageBytes = ((2 * kEncAHalf + kEncBBytes + 1023) / 1024) * 1024;
constexpr int kEncThreads = 192;
constexpr float kEncWScale = 256.0f;

struct EncTc {
  const uint8_t* w0img = nullptr;  // [n_kblocks][kEncBBytes]
  const float* b0 = nullptr;       // [128]
  int n_kblocks = 0;
  size_t smem_bytes = 0;
  bool ready = false;
};

__device__ __forceinline__ float log1p_fast(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(1.0f + x));
  return y * 0.6931471805599453f;
}
// (hi, lo) f16 split of a pair (see split2)
__device__ __forceinline__ void split2v(float2 a, uint32_t& hi, uint32_t& lo) { split2(a.x, a.y, hi, lo); }

// log1p of a pair through MUFU.LG2, packed add / multiply
__device__ __forceinline__ float2 log1p_fast2(float2 x) {
  const float2 t = __fadd2_rn(x, make_float2(1.0f, 1.0f));
  float a, b;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(t.x));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(t.y));
  return __fmul2_rn(make_float2(a, b), make_float2(0.6931471805599453f, 0.6931471805599453f));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kEncThreads, 2)
k_enc_gemm_tc(const float* __restrict__ X, int64_t ldx, int64_t n_cells, int G, EncTc enc, float* __restrict__ H1,
              int x_vec_ok) {
  // no-swizzle UMMA operands and bulk copies need 16-byte alignment only; keeping the plain array (no integer
  // re-alignment) lets the compiler see the shared address space (LDS/STS instead of generic LD/ST)
  extern __shared__ __align__(1024) uint8_t smem_tc[];
  uint8_t* const smem = smem_tc;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kEncStages * kEncStageBytes);
  uint64_t* full = bars;                    // [stages] producers (128) + B expect_tx (1)
  uint64_t* empty = bars + kEncStages;      // [stages] tcgen05.commit
  uint64_t* acc_bar = bars + 2 * kEncStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kEncStages + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb = enc.n_kblocks;
  const int64_t cell0 = (int64_t)blockIdx.x * 128;

  if (tid == 0) {
    for (int s = 0; s < kEncStages; ++s) { mbar_init(full + s, 129); mbar_init(empty + s, 1); }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------ producers
    const int rsub = lane >> 2, chunk = lane & 3;
    float4 cur[8], nxt[8];
    auto load = [&](int kb, float4 (&dst)[8]) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int row = warp * 32 + p * 8 + rsub;
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
    load(0, cur);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % kEncStages;
      if (kb + 1 < nkb) load(kb + 1, nxt);
      if (kb >= kEncStages) mbar_wait(empty + s, ((kb / kEncStages) - 1) & 1);
      uint8_t* stage = smem + s * kEncStageBytes;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int row = warp * 32 + p * 8 + rsub;
        const float4 a = cur[2 * p], b = cur[2 * p + 1];
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
#pragma unroll
      for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
    }
    // ------------------------------------------------------------ epilogue: TMEM -> H1 (+ bias, unscale)
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const int row = tid;  // == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int64_t cell = cell0 + row;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v[32];
      tmem_ld32(t_lane + c * 32, v);
      tmem_ld_wait();
      if (cell < n_cells) {
        float4* dst = reinterpret_cast<float4*>(H1 + cell * 128 + c * 32);
        const float4* bb = reinterpret_cast<const float4*>(enc.b0 + c * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 b4 = __ldg(bb + q);
          dst[q] = make_float4(fmaf(v[4 * q], 1.0f / kEncWScale, b4.x), fmaf(v[4 * q + 1], 1.0f / kEncWScale, b4.y),
                               fmaf(v[4 * q + 2], 1.0f / kEncWScale, b4.z), fmaf(v[4 * q + 3], 1.0f / kEncWScale, b4.w));
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(128);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kEncStages;
        mbar_wait(full + s, (kb / kEncStages) & 1);
        tc_fence_after();
        const uint32_t a_h = smem_u32(smem + s * kEncStageBytes), a_l = a_h + kEncAHalf;
        const uint32_t b_h = a_h + 2 * kEncAHalf, b_l = b_h + 4 * 2048;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t ah = umma_desc(a_h + ks * 2 * kEncALbo, kEncALbo, 128);
          const uint64_t al = umma_desc(a_l + ks * 2 * kEncALbo, kEncALbo, 128);
          const uint64_t bh = umma_desc(b_h + ks * 2 * 2048, 2048, 128);
          const uint64_t bl = umma_desc(b_l + ks * 2 * 2048, 2048, 128);
          umma_f16(tmem_base, ah, bh, idesc, (kb | ks) ? 1u : 0u);
          umma_f16(tmem_base, al, bh, idesc, 1u);
          umma_f16(tmem_base, ah, bl, idesc, 1u);
        }
        umma_commit(empty + s);
      }
      umma_commit(acc_bar);
    }
  } else {
    // ------------------------------------------------------------ W0 loader (TMA bulk copies)
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kEncStages;
        if (kb >= kEncStages) mbar_wait(empty + s, ((kb / kEncStages) - 1) & 1);
        mbar_expect_tx(full + s, kEncBBytes);
        bulk_g2s(smem + s * kEncStageBytes + 2 * kEncAHalf, enc.w0img + (size_t)kb * kEncBBytes, kEncBBytes, full + s);
      }
    }
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 128);
}

// pack W0 [G x 128] (+ scale 2^8) into per-K-block (hi | lo) canonical f16 images
inline int enc_tc_build(EncTc& enc, const NetW& net, const tc_detail::PMap& pm, DeviceArena& arena) {
  if (net.H != 128) return 0;  // fp32 encoder is used for other widths
  const int G = net.G;
  const float* W = tc_detail::P(pm, "qu/Dense_0/kernel");
  const int nkb = (G + kEncBK - 1) / kEncBK;
  std::vector<uint8_t> img((size_t)nkb * kEncBBytes, 0);
  for (int k = 0; k < G; ++k) {
    const int kb = k / kEncBK, kk = k % kEncBK;
    __half* hi = reinterpret_cast<__half*>(img.data() + (size_t)kb * kEncBBytes);
    __half* lo = hi + 4 * 128 * 8;
    for (int n = 0; n < 128; ++n) {
      const float w = W[(size_t)k * 128 + n] * kEncWScale;
      const __half h = __float2half_rn(w);
      const size_t e = (size_t)(kk / 8) * (128 * 8) + (size_t)n * 8 + (kk % 8);
      hi[e] = h;
      lo[e] = __float2half_rn(w - __half2float(h));
    }
  }
  int err;
  if ((err = tc_upload(arena, img.data(), img.size(), (const void**)&enc.w0img))) return err;
  enc.b0 = net.enc0.b;
  enc.n_kblocks = nkb;
  enc.smem_bytes = (size_t)kEncStages * kEncStageBytes + 256 + 1024;
  if (cudaFuncSetAttribute(k_enc_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin) != cudaSuccess) {
    g_tc_error = "cudaFuncSetAttribute(k_enc_gemm_tc) failed";
    return MRVI_ERR_CUDA;
  }
  enc.ready = true;
  return 0;
}

inline void enc_tc_launch(const EncTc& enc, const float* X, int64_t ldx, int64_t n, int G, float* H1, cudaStream_t st) {
  const int grid = (int)((n + 127) / 128);
  const int vec_ok = ((ldx % 4) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  k_enc_gemm_tc<<<grid, kEncThreads, enc.smem_bytes, st>>>(X, ldx, n, G, enc, H1, vec_ok);
}

}  // namespace mrvi
