// fp32-FFMA kernels of the local-sample-distance path (MRVI_PRECISION_FP32; parity bar 1e-5).
//
// Stage 1  k_encoder_fp32   x -> u, LayerNorm(u), query tokens, z_base     (reference _module.py:173-202, 142, 166-170)
// Stage 2  k_qz_fp32        (cell, s) -> z[cell, s, :]                     (reference _module.py:131-170, _components.py:181-230)
// Stage 3  k_dist_*         z[cell] -> D[cell] (S x S), written per cell and/or summed per group
//                                                                        (reference _model.py:545-560, 469-484)
// All arithmetic is fp32 with IEEE division / sqrt and libm-accurate tanhf / expf / log1pf.
#pragma once
#include "common.cuh"

namespace mrvi {

constexpr int kThreads = 256;
constexpr int kTM = 64;  // rows (cells or (cell,s) pairs) per CTA tile

// ------------------------------------------------------------------------------------------
// CTA-wide dense layer: out[r][n] = sum_k A[r][k] W[k][n] + b[n], r < TM, n < N (N % 4 == 0).
// A lives in shared memory (row stride lda, zero padded to Kpad), W/b in global memory (L1/L2
// resident, float4 rows).  Columns are processed in passes of <= 128; inside a pass a thread owns
// RB consecutive rows x 4 consecutive columns.  No __syncthreads inside.
// ------------------------------------------------------------------------------------------
template <int RB>
__device__ __forceinline__ void cta_dense_pass(const float* __restrict__ A, int lda, int Kpad,
                                               const float* __restrict__ W, int ldw,
                                               const float* __restrict__ bias, int ncols,
                                               float* __restrict__ out, int ldo, int TM) {
  const int ncg = ncols >> 2;
  const int nrg = kThreads / ncg;
  const int cg = threadIdx.x % ncg;
  const int rg = threadIdx.x / ncg;
  if (rg >= nrg) return;
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias != nullptr) bv = __ldg(reinterpret_cast<const float4*>(bias) + cg);
  for (int r0 = rg * RB; r0 < TM; r0 += nrg * RB) {
    float acc[RB][4];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      acc[i][0] = bv.x; acc[i][1] = bv.y; acc[i][2] = bv.z; acc[i][3] = bv.w;
    }
    for (int k = 0; k < Kpad; k += 4) {
      float4 w0 = __ldg(reinterpret_cast<const float4*>(W + (size_t)(k + 0) * ldw) + cg);
      float4 w1 = __ldg(reinterpret_cast<const float4*>(W + (size_t)(k + 1) * ldw) + cg);
      float4 w2 = __ldg(reinterpret_cast<const float4*>(W + (size_t)(k + 2) * ldw) + cg);
      float4 w3 = __ldg(reinterpret_cast<const float4*>(W + (size_t)(k + 3) * ldw) + cg);
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        const int r = r0 + i;
        float4 a = (r < TM) ? *reinterpret_cast<const float4*>(A + (size_t)r * lda + k)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
        acc[i][0] = fmaf(a.x, w0.x, acc[i][0]); acc[i][1] = fmaf(a.x, w0.y, acc[i][1]);
        acc[i][2] = fmaf(a.x, w0.z, acc[i][2]); acc[i][3] = fmaf(a.x, w0.w, acc[i][3]);
        acc[i][0] = fmaf(a.y, w1.x, acc[i][0]); acc[i][1] = fmaf(a.y, w1.y, acc[i][1]);
        acc[i][2] = fmaf(a.y, w1.z, acc[i][2]); acc[i][3] = fmaf(a.y, w1.w, acc[i][3]);
        acc[i][0] = fmaf(a.z, w2.x, acc[i][0]); acc[i][1] = fmaf(a.z, w2.y, acc[i][1]);
        acc[i][2] = fmaf(a.z, w2.z, acc[i][2]); acc[i][3] = fmaf(a.z, w2.w, acc[i][3]);
        acc[i][0] = fmaf(a.w, w3.x, acc[i][0]); acc[i][1] = fmaf(a.w, w3.y, acc[i][1]);
        acc[i][2] = fmaf(a.w, w3.z, acc[i][2]); acc[i][3] = fmaf(a.w, w3.w, acc[i][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const int r = r0 + i;
      if (r < TM)
        *reinterpret_cast<float4*>(out + (size_t)r * ldo + 4 * cg) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
  }
}

// A: smem [TM][lda]; out: smem [TM][ldo] (columns [0, d.ld)).  lda, ldo multiples of 4.
__device__ __forceinline__ void cta_dense(const float* A, int lda, const DenseW& d, float* out,
                                          int ldo, int TM) {
  for (int c0 = 0; c0 < d.ld; c0 += 128) {
    const int ncols = min(128, d.ld - c0);
    const int ncg = ncols >> 2;
    const int per = (TM * ncg + kThreads - 1) / kThreads;  // rows per thread if spread evenly
    const float* W = d.w + c0;
    const float* b = d.b ? d.b + c0 : nullptr;
    if (per >= 8) cta_dense_pass<8>(A, lda, d.Kpad, W, d.ld, b, ncols, out + c0, ldo, TM);
    else if (per >= 4) cta_dense_pass<4>(A, lda, d.Kpad, W, d.ld, b, ncols, out + c0, ldo, TM);
    else if (per >= 2) cta_dense_pass<2>(A, lda, d.Kpad, W, d.ld, b, ncols, out + c0, ldo, TM);
    else cta_dense_pass<1>(A, lda, d.Kpad, W, d.ld, b, ncols, out + c0, ldo, TM);
  }
}

// ------------------------------------------------------------------------------------------
// Row LayerNorm (+ affine, + activation, + residual): one warp per row.
//   y = (v - mean) * (rsqrt(var + eps) * scale[c]) + bias[c];  var = max(0, E[v^2] - mean^2)
//   ACT: 0 none, 1 relu, 2 gelu(tanh).  Then optional  y += resid[r][c].
// scale/bias may be per-row conditional tables (row_scale/row_bias point at the row's vector).
// Columns >= N of `out` (up to Npad) are zeroed so the next dense layer may read padded K.
// ------------------------------------------------------------------------------------------
template <int ACT>
__device__ __forceinline__ void cta_ln_rows(const float* in, int ldi, int TM, int N, int Npad,
                                            const float* __restrict__ scale,
                                            const float* __restrict__ bias,
                                            const float* const* row_scale,  // smem [TM] or null
                                            const float* const* row_bias,   // smem [TM] or null
                                            const float* const* row_add,    // smem [TM] or null (added after ACT)
                                            const float* resid, int ldr, float* out, int ldo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float invN = 1.0f / (float)N;
  for (int r = warp; r < TM; r += kThreads / 32) {
    const float* v = in + (size_t)r * ldi;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < N; c += 32) {
      float x = v[c];
      s1 += x;
      s2 = fmaf(x, x, s2);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    const float mean = s1 * invN;
    const float var = fmaxf(s2 * invN - mean * mean, 0.f);
    const float rstd = 1.0f / sqrtf(var + kLnEps);
    const float* sc = row_scale ? row_scale[r] : scale;
    const float* bi = row_bias ? row_bias[r] : bias;
    const float* ad = row_add ? row_add[r] : nullptr;
    for (int c = lane; c < Npad; c += 32) {
      float y = 0.f;
      if (c < N) {
        float mul = rstd;
        if (sc) mul *= __ldg(sc + c);
        y = (v[c] - mean) * mul;
        if (bi) y += __ldg(bi + c);
        if (ACT == 1) y = fmaxf(y, 0.f);
        if (ACT == 2) y = gelu_tanh_f(y);
        if (ad) y += __ldg(ad + c);
        if (resid) y += resid[(size_t)r * ldr + c];
      }
      out[(size_t)r * ldo + c] = y;
    }
  }
}

// ResnetBlock on a tile: x [TM][ldx] (n_in cols, zero padded) -> y [TM][ldy] (n_out cols, zero padded).
// hid: scratch [TM][ldh] with ldh >= 256 when the block has a skip Dense, else >= 128.
template <int ACT>
__device__ __forceinline__ void cta_resnet_block(const ResnetW& rb, const float* x, int ldx,
                                                 float* hid, int ldh, float* y, int ldy, int TM) {
  cta_dense(x, ldx, rb.d0, hid, ldh, TM);
  if (rb.dskip.w) cta_dense(x, ldx, rb.dskip, hid + kResnetHidden, ldh, TM);
  __syncthreads();
  // t = act(LN(h)) + skip  (in place over hid[:, :128])
  cta_ln_rows<ACT>(hid, ldh, TM, kResnetHidden, kResnetHidden, rb.ln0.scale, rb.ln0.bias, nullptr,
                   nullptr, nullptr, rb.dskip.w ? hid + kResnetHidden : x, rb.dskip.w ? ldh : ldx,
                   hid, ldh);
  __syncthreads();
  cta_dense(hid, ldh, rb.dout, y, ldy, TM);
  __syncthreads();
  cta_ln_rows<ACT>(y, ldy, TM, rb.dout.N, rb.dout.ld, rb.ln1.scale, rb.ln1.bias, nullptr, nullptr,
                   nullptr, nullptr, 0, y, ldy);
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Stage 1: encoder.  One CTA = kTM cells.
// smem layout (floats): xs[kTM][36] | ha[kTM][ldh] | hb[kTM][ldh] | row pointer tables
// ------------------------------------------------------------------------------------------
struct EncSmem {
  int ldh;        // >= max(H, 128) + 4, multiple of 4; >= 260 if any block has a skip dense
  int hid_ld;     // scratch width for resnet hidden (128 or 256) + 4
};

__global__ void __launch_bounds__(kThreads)
k_encoder_fp32(const __grid_constant__ NetW net, const float* __restrict__ X, int64_t ldx,
               const int32_t* __restrict__ sample_idx, int64_t n_cells, CellBuf cb, int ldh,
               int ldhid, int x_vec_ok, const float* __restrict__ h1_in) {
  extern __shared__ __align__(16) float smem[];
  const int H = net.H;
  float* xs = smem;                       // [kTM][36]
  float* ha = xs + kTM * 36;              // [kTM][ldh]
  float* hb = ha + (size_t)kTM * ldh;     // [kTM][ldh]
  float* hid = hb + (size_t)kTM * ldh;    // [kTM][ldhid]
  const float** rp0 = reinterpret_cast<const float**>(hid + (size_t)kTM * ldhid);  // [3][kTM]
  const float** rp1 = rp0 + kTM;
  const float** rp2 = rp1 + kTM;
  int* sid_s = reinterpret_cast<int*>(rp2 + kTM);

  const int64_t cell0 = (int64_t)blockIdx.x * kTM;
  const int TM = (int)min((int64_t)kTM, n_cells - cell0);
  const int tid = threadIdx.x;
  if (tid < kTM) {
    int s = (tid < TM) ? sample_idx[cell0 + tid] : 0;
    s = min(max(s, 0), net.S - 1);
    sid_s[tid] = s;
  }

  // ---- Dense_0 over K = G in chunks of 32, accumulators in registers (8 rows x 4 cols / thread,
  //      column passes of 128)
  const int Hld = net.enc0.ld;  // roundup(H,4)
  if (h1_in != nullptr) {
    // Dense_0 (+ bias) was computed by the tensor-core GEMM (k_enc_gemm_tc): load this tile's rows
    for (int idx = tid; idx < kTM * (Hld / 4); idx += kThreads) {
      const int r = idx / (Hld / 4), c4 = idx % (Hld / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < TM) v = __ldg(reinterpret_cast<const float4*>(h1_in + (cell0 + r) * Hld) + c4);
      *reinterpret_cast<float4*>(ha + (size_t)r * ldh + 4 * c4) = v;
    }
  }
  for (int c0 = 0; h1_in == nullptr && c0 < Hld; c0 += 128) {
    const int ncols = min(128, Hld - c0);
    const int ncg = ncols >> 2;
    const int nrg = kThreads / ncg;
    const int cg = tid % ncg, rg = tid / ncg;
    const bool active = rg < nrg;
    // rows owned: r = rg + nrg * i  (i < RBX)
    constexpr int RBX = 8;
    float acc[RBX][4];
    float4 bv = active ? __ldg(reinterpret_cast<const float4*>(net.enc0.b + c0) + cg)
                       : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < RBX; ++i) { acc[i][0] = bv.x; acc[i][1] = bv.y; acc[i][2] = bv.z; acc[i][3] = bv.w; }
    const int row_iters = (kTM + nrg * RBX - 1) / (nrg * RBX);  // 1 when ncg == 32
    for (int it = 0; it < row_iters; ++it) {
      if (it > 0) {
#pragma unroll
        for (int i = 0; i < RBX; ++i) { acc[i][0] = bv.x; acc[i][1] = bv.y; acc[i][2] = bv.z; acc[i][3] = bv.w; }
      }
      for (int k0 = 0; k0 < net.enc0.Kpad; k0 += 32) {
        __syncthreads();
        {  // load + log1p a [kTM][32] chunk of X: thread -> row tid/4, 8 columns
          const int r = tid >> 2, cq = (tid & 3) * 8;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = 0.f;
          if (r < TM) {
            const float* src = X + (cell0 + r) * ldx + k0 + cq;
            if (x_vec_ok && k0 + cq + 8 <= net.G) {
              float4 a = __ldg(reinterpret_cast<const float4*>(src));
              float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
              v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) if (k0 + cq + j < net.G) v[j] = __ldg(src + j);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) xs[r * 36 + cq + j] = log1pf(v[j]);
        }
        __syncthreads();
        if (active) {
          const int kmax = min(32, net.enc0.Kpad - k0);
          for (int kk = 0; kk < kmax; kk += 4) {
            const float* Wp = net.enc0.w + (size_t)(k0 + kk) * Hld + c0;
            float4 w0 = __ldg(reinterpret_cast<const float4*>(Wp) + cg);
            float4 w1 = __ldg(reinterpret_cast<const float4*>(Wp + Hld) + cg);
            float4 w2 = __ldg(reinterpret_cast<const float4*>(Wp + 2 * Hld) + cg);
            float4 w3 = __ldg(reinterpret_cast<const float4*>(Wp + 3 * Hld) + cg);
#pragma unroll
            for (int i = 0; i < RBX; ++i) {
              const int r = (it * RBX + i) * nrg + rg;
              if (r < kTM) {
                float4 a = *reinterpret_cast<const float4*>(xs + r * 36 + kk);
                acc[i][0] = fmaf(a.x, w0.x, acc[i][0]); acc[i][1] = fmaf(a.x, w0.y, acc[i][1]);
                acc[i][2] = fmaf(a.x, w0.z, acc[i][2]); acc[i][3] = fmaf(a.x, w0.w, acc[i][3]);
                acc[i][0] = fmaf(a.y, w1.x, acc[i][0]); acc[i][1] = fmaf(a.y, w1.y, acc[i][1]);
                acc[i][2] = fmaf(a.y, w1.z, acc[i][2]); acc[i][3] = fmaf(a.y, w1.w, acc[i][3]);
                acc[i][0] = fmaf(a.z, w2.x, acc[i][0]); acc[i][1] = fmaf(a.z, w2.y, acc[i][1]);
                acc[i][2] = fmaf(a.z, w2.z, acc[i][2]); acc[i][3] = fmaf(a.z, w2.w, acc[i][3]);
                acc[i][0] = fmaf(a.w, w3.x, acc[i][0]); acc[i][1] = fmaf(a.w, w3.y, acc[i][1]);
                acc[i][2] = fmaf(a.w, w3.z, acc[i][2]); acc[i][3] = fmaf(a.w, w3.w, acc[i][3]);
              }
            }
          }
        }
      }
      if (active) {
#pragma unroll
        for (int i = 0; i < RBX; ++i) {
          const int r = (it * RBX + i) * nrg + rg;
          if (r < kTM)
            *reinterpret_cast<float4*>(ha + (size_t)r * ldh + c0 + 4 * cg) =
                make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
      }
    }
  }
  // per-row conditional tables for ConditionalNormalization_0
  if (tid < kTM) {
    rp0[tid] = net.gamma[0] + (size_t)sid_s[tid] * H;
    rp1[tid] = net.beta[0] + (size_t)sid_s[tid] * H;
  }
  __syncthreads();
  // ConditionalNormalization_0 + gelu  (reference _module.py:190-195, _components.py:142-160)
  cta_ln_rows<2>(ha, ldh, kTM, H, Hld, nullptr, nullptr, rp0, rp1, nullptr, nullptr, 0, ha, ldh);
  __syncthreads();
  // Dense_1
  cta_dense(ha, ldh, net.enc1, hb, ldh, kTM);
  if (tid < kTM) {
    rp0[tid] = net.gamma[1] + (size_t)sid_s[tid] * H;
    rp1[tid] = net.beta[1] + (size_t)sid_s[tid] * H;
    rp2[tid] = net.sample_effect + (size_t)sid_s[tid] * H;
  }
  __syncthreads();
  // ConditionalNormalization_1 + gelu, then + sample_effect[sid]  (reference _module.py:196-199)
  cta_ln_rows<2>(hb, ldh, kTM, H, Hld, nullptr, nullptr, rp0, rp1, rp2, nullptr, 0, hb, ldh);
  __syncthreads();
  // NormalDistOutputNN: ResnetBlocks (relu/relu) then the mean head (reference _components.py:92-95)
  float* cur = hb;
  float* nxt = ha;
  for (int b = 0; b < net.n_enc_blocks; ++b) {
    cta_resnet_block<1>(net.enc_blocks[b], cur, ldh, hid, ldhid, nxt, ldh, kTM);
    float* t = cur; cur = nxt; nxt = t;
  }
  // mean head -> u (into nxt[:, 0:Lq4])
  cta_dense(cur, ldh, net.enc_mean, nxt, ldh, kTM);
  __syncthreads();
  const int Lq = net.Lq, Lq4 = net.enc_mean.ld;
  if (cb.u_scale != nullptr && net.enc_scale.w != nullptr) {
    // scale head: softplus(Dense(h)) + 1e-5  (reference _components.py:96-97), only for the sampled paths
    cta_dense(cur, ldh, net.enc_scale, hid, ldhid, kTM);
    __syncthreads();
    for (int idx = tid; idx < TM * Lq; idx += kThreads) {
      const int r = idx / Lq, c = idx % Lq;
      const float v = hid[(size_t)r * ldhid + c];
      cb.u_scale[(cell0 + r) * Lq + c] = (v > 20.f ? v : log1pf(expf(v))) + 1e-5f;
    }
    __syncthreads();
  }
  // u_ln = LayerNorm(u; scale, bias) -> cur[:, 0:Lq4]
  cta_ln_rows<0>(nxt, ldh, kTM, Lq, Lq4, net.u_ln.scale, net.u_ln.bias, nullptr, nullptr, nullptr,
                 nullptr, 0, cur, ldh);
  __syncthreads();
  // q tokens = u_ln @ Wq_tok -> hid[:, 0:E4];  z_base = u @ Wzb + b -> hid[:, 64: 64+L4]
  cta_dense(cur, ldh, net.q_tok, hid, ldhid, kTM);
  if (net.has_zbase) cta_dense(nxt, ldh, net.zbase, hid + 64, ldhid, kTM);
  __syncthreads();
  for (int idx = tid; idx < TM * Lq; idx += kThreads) {
    const int r = idx / Lq, c = idx % Lq;
    cb.u[(cell0 + r) * Lq + c] = nxt[(size_t)r * ldh + c];
    cb.u_ln[(cell0 + r) * Lq + c] = cur[(size_t)r * ldh + c];
  }
  for (int idx = tid; idx < TM * net.E; idx += kThreads) {
    const int r = idx / net.E, c = idx % net.E;
    cb.q_tok[(cell0 + r) * net.E + c] = hid[(size_t)r * ldhid + c];
  }
  for (int idx = tid; idx < TM * net.L; idx += kThreads) {
    const int r = idx / net.L, c = idx % net.L;
    cb.zbase[(cell0 + r) * net.L + c] =
        net.has_zbase ? hid[(size_t)r * ldhid + 64 + c] : nxt[(size_t)r * ldh + c];
  }
}

// ------------------------------------------------------------------------------------------
// Stage 2: q(z | u, s) for kTM consecutive (cell, s) rows per CTA.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_qz_fp32(const __grid_constant__ NetW net, CellBuf cb, int64_t n_cells, int rows_per_cell, float* __restrict__ z_out,
          int lda_att, int ldhid, int ldm, int ldc) {
  extern __shared__ __align__(16) float smem[];
  float* att = smem;                                  // [kTM][lda_att]  attention output (E*C cols)
  float* hid = att + (size_t)kTM * lda_att;           // [kTM][ldhid]    resnet hidden (+ skip)
  float* mbuf = hid + (size_t)kTM * ldhid;            // [kTM][ldm]      M-wide block outputs
  float* cbuf = mbuf + (size_t)kTM * ldm;             // [kTM][ldc]      concat [u_ln, eps_] / final output
  // rows_per_cell = S on the mean path (row -> (cell, s)); the sampled paths re-map rows through cb.per_row /
  // cb.cell_sample (see CellBuf)
  const int S = rows_per_cell, E = net.E, C = net.C, nh = net.nh, hd = net.C, Lq = net.Lq, L = net.L;
  const int64_t row0 = (int64_t)blockIdx.x * kTM;
  const int64_t n_rows = n_cells * S;
  const int tid = threadIdx.x;

  // ---- attention: 4 threads per row, each owns query tokens i = q4, q4+4, ...   (reference _components.py:195-210)
  {
    const int r = tid >> 2, q4 = tid & 3;
    const int64_t row = row0 + r;
    const bool valid = row < n_rows;
    const int64_t cell = valid ? row / S : 0;
    const int s = valid ? (cb.cell_sample ? cb.cell_sample[cell] : (int)(row % S)) : 0;
    const int64_t fi = cb.per_row ? (valid ? row : 0) : cell;
    const float* ktab = net.ktab + (size_t)s * E * nh * hd;
    const float* vtab = net.vtab + (size_t)s * E * nh * hd;
    const float inv_sqrt_d = 1.0f / sqrtf((float)hd);
    for (int i = q4; i < E; i += 4) {
      const float qt = valid ? cb.q_tok[fi * E + i] : 0.f;
      float o[kMaxHeadDim];
#pragma unroll
      for (int c = 0; c < kMaxHeadDim; ++c) o[c] = (c < C) ? __ldg(net.bo + c) : 0.f;
      for (int h = 0; h < nh; ++h) {
        float qv[kMaxHeadDim];
#pragma unroll
        for (int d = 0; d < kMaxHeadDim; ++d)
          qv[d] = (d < hd) ? (qt * __ldg(net.wq + h * hd + d) + __ldg(net.bq + h * hd + d)) * inv_sqrt_d : 0.f;
        float sc[kMaxE];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kMaxE; ++j) {
          if (j < E) {
            const float* kr = ktab + ((size_t)j * nh + h) * hd;
            float a = 0.f;
#pragma unroll
            for (int d = 0; d < kMaxHeadDim; ++d) if (d < hd) a = fmaf(qv[d], __ldg(kr + d), a);
            sc[j] = a;
            mx = fmaxf(mx, a);
          }
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxE; ++j) if (j < E) { sc[j] = expf(sc[j] - mx); sum += sc[j]; }
        float av[kMaxHeadDim];
#pragma unroll
        for (int d = 0; d < kMaxHeadDim; ++d) av[d] = 0.f;
#pragma unroll
        for (int j = 0; j < kMaxE; ++j) {
          if (j < E) {
            const float p = sc[j] / sum;
            const float* vr = vtab + ((size_t)j * nh + h) * hd;
#pragma unroll
            for (int d = 0; d < kMaxHeadDim; ++d) if (d < hd) av[d] = fmaf(p, __ldg(vr + d), av[d]);
          }
        }
#pragma unroll
        for (int d = 0; d < kMaxHeadDim; ++d)
          if (d < hd) {
#pragma unroll
            for (int c = 0; c < kMaxHeadDim; ++c)
              if (c < C) o[c] = fmaf(av[d], __ldg(net.wo + ((size_t)h * hd + d) * C + c), o[c]);
          }
      }
#pragma unroll
      for (int c = 0; c < kMaxHeadDim; ++c) if (c < C) att[(size_t)r * lda_att + i * C + c] = o[c];
    }
    // zero the K padding of the attention buffer
    for (int c = E * C + q4; c < lda_att; c += 4) att[(size_t)r * lda_att + c] = 0.f;
  }
  __syncthreads();
  // ---- MLP_0: ResnetBlock(E*C -> 128 -> M, gelu) then Dense(M -> E)     (reference _components.py:216-221)
  cta_resnet_block<2>(net.mlp0_block, att, lda_att, hid, ldhid, mbuf, ldm, kTM);
  // eps_ -> cbuf[:, Lq : Lq+E]; u_ln -> cbuf[:, 0:Lq]                    (reference _components.py:222)
  {
    // dense M -> E written at column offset Lq requires 4-aligned offset; write to att scratch first
    cta_dense(mbuf, ldm, net.mlp0_out, att, lda_att, kTM);
    __syncthreads();
    const int W = Lq + E;
    for (int idx = tid; idx < kTM * ldc; idx += kThreads) {
      const int r = idx / ldc, c = idx % ldc;
      const int64_t row = row0 + r;
      float v = 0.f;
      if (row < n_rows) {
        if (c < Lq) v = cb.u_ln[(cb.per_row ? row : row / S) * Lq + c];
        else if (c < W) v = att[(size_t)r * lda_att + (c - Lq)];
      }
      cbuf[(size_t)r * ldc + c] = v;
    }
    __syncthreads();
  }
  // ---- MLP_1: n_qz_blocks ResnetBlocks then Dense(M -> out_dim)         (reference _components.py:223-229)
  {
    const float* x = cbuf;
    int ldx_ = ldc;
    float* y = mbuf;
    for (int b = 0; b < net.n_qz_blocks; ++b) {
      // alternate between mbuf and att as block outputs (both >= M+pad wide)
      y = (b & 1) ? att : mbuf;
      const int ldy = (b & 1) ? lda_att : ldm;
      cta_resnet_block<2>(net.mlp1_blocks[b], x, ldx_, hid, ldhid, y, ldy, kTM);
      x = y; ldx_ = ldy;
    }
    cta_dense(x, ldx_, net.mlp1_out, hid, ldhid, kTM);
    __syncthreads();
  }
  // ---- z = z_base + eps[:L]                                              (reference _module.py:326-331)
  for (int idx = tid; idx < kTM * L; idx += kThreads) {
    const int r = idx / L, c = idx % L;
    const int64_t row = row0 + r;
    if (row < n_rows) {
      z_out[row * L + c] = cb.zbase[(cb.per_row ? row : row / S) * L + c] + hid[(size_t)r * ldhid + c];
      if (cb.raw_out != nullptr && net.out_dim == 2 * L) cb.raw_out[row * L + c] = hid[(size_t)r * ldhid + L + c];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Stage 3: pairwise distances.  NORM: 0 l2, 1 l1, 2 linf.
// Work item = sorted position i (cell = order ? order[i] : i).  A worker (CTA or warp) owns a
// contiguous range of positions, keeps per-thread accumulators for fixed (a, b) entries over a
// run of equal composite group code, and flushes them with float64 atomics when the code changes.
// ------------------------------------------------------------------------------------------
template <int NORM>
__device__ __forceinline__ float dist_accum(float acc, float d) {
  if (NORM == 0) return fmaf(d, d, acc);
  if (NORM == 1) return acc + fabsf(d);
  return fmaxf(acc, fabsf(d));
}
template <int NORM>
__device__ __forceinline__ float dist_final(float acc) { return NORM == 0 ? sqrtf(acc) : acc; }

__device__ __forceinline__ void flush_entry(const GroupPlan& gp, int64_t cell, int S, int a, int b, float v) {
  for (int k = 0; k < gp.n_keys; ++k) {
    const int g = gp.codes[k][cell];
    if (g >= 0 && g < gp.n_groups[k]) atomicAdd(gp.sums[k] + ((size_t)g * S + a) * S + b, (double)v);
  }
}

// Small S (S*S <= 1024): one warp per cell, lane owns entries e = lane + 32 q.
template <int NORM>
__global__ void __launch_bounds__(kThreads)
k_dist_small(const float* __restrict__ z, int64_t n_cells, int S, int L, float* __restrict__ cell_out,
             GroupPlan gp, int cells_per_warp) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Lp = L | 1;  // odd stride: conflict-free column reads
  float* zs = smem + (size_t)warp * S * Lp;
  const int64_t wid = (int64_t)blockIdx.x * (kThreads / 32) + warp;
  const int64_t p0 = wid * cells_per_warp;
  const int64_t p1 = min(n_cells, p0 + cells_per_warp);
  const int SS = S * S;
  constexpr int QMAX = 32;
  float acc[QMAX];
#pragma unroll
  for (int q = 0; q < QMAX; ++q) acc[q] = 0.f;
  int cur_code = -1;
  int64_t cur_cell = -1;
  for (int64_t p = p0; p < p1; ++p) {
    const int64_t cell = gp.order ? gp.order[p] : p;
    const int code = gp.n_keys ? gp.comp_sorted[p] : 0;
    if (gp.n_keys && cur_cell >= 0 && code != cur_code) {
#pragma unroll
      for (int q = 0; q < QMAX; ++q) {
        const int e = lane + 32 * q;
        if (e < SS) { flush_entry(gp, cur_cell, S, e / S, e % S, acc[q]); acc[q] = 0.f; }
      }
    }
    cur_code = code; cur_cell = cell;
    __syncwarp();
    for (int i = lane; i < S * L; i += 32) zs[(i / L) * Lp + (i % L)] = z[(size_t)cell * S * L + i];
    __syncwarp();
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
      const int e = lane + 32 * q;
      if (e < SS) {
        const int a = e / S, b = e % S;
        float d = 0.f;
        for (int l = 0; l < L; ++l) d = dist_accum<NORM>(d, zs[b * Lp + l] - zs[a * Lp + l]);
        d = dist_final<NORM>(d);
        if (cell_out) cell_out[(size_t)cell * SS + e] = d;
        acc[q] += d;
      }
    }
  }
  if (gp.n_keys && cur_cell >= 0) {
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
      const int e = lane + 32 * q;
      if (e < SS) flush_entry(gp, cur_cell, S, e / S, e % S, acc[q]);
    }
  }
}

// General S: CTA = (range of sorted positions) x (64 x 64 block of the S x S matrix); thread owns a
// 4 x 4 register tile: rows ty*4.., cols tx*4.. (tx fastest -> float4 row stores).
constexpr int kDB = 64;
template <int NORM>
__global__ void __launch_bounds__(kThreads)
k_dist_tiled(const float* __restrict__ z, int64_t n_cells, int S, int L, float* __restrict__ cell_out,
             GroupPlan gp, int cells_per_cta) {
  extern __shared__ __align__(16) float smem[];
  const int Lp = (L + 3) & ~3;
  float* za = smem;                     // [kDB][Lp]   rows of block I (row-major, broadcast reads)
  float* zbT = za + kDB * Lp;           // [Lp][kDB]   rows of block J, transposed (float4 column reads)
  const int nb = (S + kDB - 1) / kDB;
  const int bi = blockIdx.y / nb, bj = blockIdx.y % nb;
  const int i0 = bi * kDB, j0 = bj * kDB;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t p0 = (int64_t)blockIdx.x * cells_per_cta;
  const int64_t p1 = min(n_cells, p0 + cells_per_cta);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  int cur_code = -1;
  int64_t cur_cell = -1;
  const bool vec_ok = (S % 4) == 0;
  for (int64_t p = p0; p < p1; ++p) {
    const int64_t cell = gp.order ? gp.order[p] : p;
    const int code = gp.n_keys ? gp.comp_sorted[p] : 0;
    if (gp.n_keys && cur_cell >= 0 && code != cur_code) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int i = i0 + ty * 4 + a, j = j0 + tx * 4 + b;
          if (i < S && j < S) flush_entry(gp, cur_cell, S, i, j, acc[a][b]);
          acc[a][b] = 0.f;
        }
    }
    cur_code = code; cur_cell = cell;
    __syncthreads();
    const float* zc = z + (size_t)cell * S * L;
    for (int idx = threadIdx.x; idx < kDB * Lp; idx += kThreads) {
      const int r = idx / Lp, l = idx % Lp;
      za[idx] = (i0 + r < S && l < L) ? zc[(size_t)(i0 + r) * L + l] : 0.f;
      zbT[l * kDB + r] = (j0 + r < S && l < L) ? zc[(size_t)(j0 + r) * L + l] : 0.f;
    }
    __syncthreads();
    float d[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) d[a][b] = 0.f;
    for (int l = 0; l < L; ++l) {
      const float4 bv = *reinterpret_cast<const float4*>(zbT + l * kDB + tx * 4);
      const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const float av = za[(ty * 4 + a) * Lp + l];
#pragma unroll
        for (int b = 0; b < 4; ++b) d[a][b] = dist_accum<NORM>(d[a][b], bb[b] - av);
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + ty * 4 + a;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        d[a][b] = dist_final<NORM>(d[a][b]);
        acc[a][b] += d[a][b];
      }
      if (cell_out && i < S) {
        float* dst = cell_out + ((size_t)cell * S + i) * S + j0 + tx * 4;
        if (vec_ok && j0 + tx * 4 + 3 < S) {
          *reinterpret_cast<float4*>(dst) = make_float4(d[a][0], d[a][1], d[a][2], d[a][3]);
        } else {
#pragma unroll
          for (int b = 0; b < 4; ++b) if (j0 + tx * 4 + b < S) dst[b] = d[a][b];
        }
      }
    }
  }
  if (gp.n_keys && cur_cell >= 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = i0 + ty * 4 + a, j = j0 + tx * 4 + b;
        if (i < S && j < S) flush_entry(gp, cur_cell, S, i, j, acc[a][b]);
      }
  }
}

// ------------------------------------------------------------------------------------------
// Per-sample tables, once per handle: e_s = LN(Embed[s]); kv tokens = e_s @ W_kv;
// K_s[j,h,:] = kv_j * wk[h,:] + bk[h,:]; V_s likewise.  (reference _module.py:144-147,
// _components.py:198-205).  One thread per sample.
// ------------------------------------------------------------------------------------------
__global__ void k_sample_tables(const float* __restrict__ embed, const float* __restrict__ ln_scale,
                                const float* __restrict__ ln_bias, const float* __restrict__ w_kv,
                                const float* __restrict__ wk, const float* __restrict__ bk,
                                const float* __restrict__ wv, const float* __restrict__ bv, int S, int E,
                                int nh, int hd, float* __restrict__ ktab, float* __restrict__ vtab,
                                float* __restrict__ kvtok) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float e[kMaxE];
  float s1 = 0.f, s2 = 0.f;
  for (int j = 0; j < E; ++j) { e[j] = embed[(size_t)s * E + j]; s1 += e[j]; s2 = fmaf(e[j], e[j], s2); }
  const float mean = s1 / (float)E;
  const float var = fmaxf(s2 / (float)E - mean * mean, 0.f);
  const float rstd = 1.0f / sqrtf(var + kLnEps);
  for (int j = 0; j < E; ++j) e[j] = (e[j] - mean) * (rstd * ln_scale[j]) + ln_bias[j];
  for (int j = 0; j < E; ++j) {
    float kv = 0.f;
    for (int i = 0; i < E; ++i) kv = fmaf(e[i], w_kv[(size_t)i * E + j], kv);
    kvtok[(size_t)s * E + j] = kv;
    for (int h = 0; h < nh; ++h)
      for (int d = 0; d < hd; ++d) {
        ktab[(((size_t)s * E + j) * nh + h) * hd + d] = kv * wk[h * hd + d] + bk[h * hd + d];
        vtab[(((size_t)s * E + j) * nh + h) * hd + d] = kv * wv[h * hd + d] + bv[h * hd + d];
      }
  }
}

// CSR rows [0, n) of one chunk -> dense row-major X [n][G] (the buffer is zeroed before): one warp per row.
// indptr_rel holds the chunk's n+1 row pointers (absolute), nz0 = indptr of the chunk's first row.
__global__ void k_csr_densify(const int64_t* __restrict__ indptr, int64_t nz0, const int32_t* __restrict__ indices,
                              const float* __restrict__ data, int64_t n, int G, float* __restrict__ X) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const int64_t b = indptr[row] - nz0, e = indptr[row + 1] - nz0;
  for (int64_t k = b + lane; k < e; k += 32) {
    const int c = indices[k];
    if (c >= 0 && c < G) X[row * G + c] = data[k];
  }
}

// widen a narrowed count chunk (uint8 / uint16, see host_pack.cpp) back to the identical float32 values
template <typename T>
__global__ void k_unpack_counts(const T* __restrict__ in, float* __restrict__ X, int64_t n) {
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i0 + 16 <= n && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    T v[16];
    if (sizeof(T) == 1) {
      *reinterpret_cast<uint4*>(v) = __ldg(reinterpret_cast<const uint4*>(in + i0));
    } else {
      reinterpret_cast<uint4*>(v)[0] = __ldg(reinterpret_cast<const uint4*>(in + i0));
      reinterpret_cast<uint4*>(v)[1] = __ldg(reinterpret_cast<const uint4*>(in + i0) + 1);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<float4*>(X + i0 + 4 * q) = make_float4((float)v[4 * q], (float)v[4 * q + 1], (float)v[4 * q + 2], (float)v[4 * q + 3]);
  } else {
    for (int64_t i = i0; i < min(n, i0 + 16); ++i) X[i] = (float)in[i];
  }
}

// composite group code of each cell (mixed radix over (code_k + 1)), for the per-chunk sort
__global__ void k_composite_codes(GroupPlan gp, const int32_t* n_groups_dev, int64_t n, int32_t* comp,
                                  int32_t* idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t c = 0;
  for (int k = 0; k < gp.n_keys; ++k) {
    int g = gp.codes[k][i];
    if (g < -1 || g >= n_groups_dev[k]) g = -1;
    c = c * (n_groups_dev[k] + 1) + (g + 1);
  }
  comp[i] = c;
  idx[i] = (int32_t)i;
}

}  // namespace mrvi
