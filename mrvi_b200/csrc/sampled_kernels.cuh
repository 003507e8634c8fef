// Sampled paths of the local-sample-distance computation:
//   use_mean=False            -> "sampled_distances"     (n, mc, S, S)   reference _model.py:406-433
//   normalize_distances=True  -> "normalized_distances"  (n, S, S)       reference _model.py:435-450, 508-540
//
// In the reference every vmapped counterfactual call s owns a PRNG stream and draws its own
// u = qu.rsample(rng_s, (mc,)) (reference _model.py:136-159, _module.py:315-318), so each (cell, mc, s) row of
// the q(z|u,s) chain has its own u.  Here the draws are either supplied by the caller (exact parity against the
// oracle on the same draws) or generated on the device with the counter-based Philox4x32-10 generator keyed by
// (seed; global cell, mc index, sample, stream), which makes a result independent of chunking and of how cells
// are sharded over GPUs.  JAX's threefry streams are not reproduced (they cannot be without jax): parity with
// the reference itself is statistical on this path.
#pragma once
#include "common.cuh"

namespace mrvi {

constexpr int kMaxLq = 96;

// ---------------------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// four standard normals from one Philox block (Box-Muller on 24-bit uniforms in (0, 1))
__device__ __forceinline__ float4 philox_normal4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed) {
  const uint4 x = philox4x32_10(make_uint4(c0, c1, c2, c3), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float s = 1.0f / 16777216.0f;
  const float u0 = ((float)(x.x >> 8) + 0.5f) * s, u1 = ((float)(x.y >> 8) + 0.5f) * s;
  const float u2 = ((float)(x.z >> 8) + 0.5f) * s, u3 = ((float)(x.w >> 8) + 0.5f) * s;
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0f, s1, c1f;
  sincospif(2.0f * u1, &s0, &c0f);
  sincospif(2.0f * u3, &s1, &c1f);
  return make_float4(r0 * c0f, r0 * s0, r1 * c1f, r1 * s1);
}

// ----------------------------------------------------------------------------------------
// One thread per row.  Row index = (cell * A + a) * B + b:
//   counterfactual rows: A = mc, B = S  (a = mc index, b = counterfactual sample)   noise[(b * A + a)][cell][Lq]
//   baseline rows:       A = 2 mc, B = 1 (a = draw index)                            noise[a][cell][Lq]
// (explicit noise is indexed by the call-local cell noise_cell0 + cell of noise_n; Philox by the global cell gcell0 + cell)
// u = mean[cell] + scale[cell] * noise (reference _module.py:315-318), then the per-row inputs of the chain:
// LayerNorm(u) (reference _module.py:142), query tokens (reference _components.py:195-197), z_base (reference _module.py:166-170).
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_sample_rows(const __grid_constant__ NetW net, const float* __restrict__ mean, const float* __restrict__ scale,
              int64_t n_cells, int A, int B, const float* __restrict__ noise, int64_t noise_cell0, int64_t noise_n,
              int64_t gcell0, uint64_t seed, uint32_t stream, float* __restrict__ u_ln_out, float* __restrict__ q_tok_out,
              float* __restrict__ zbase_out, float* __restrict__ u_draw_out = nullptr) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_cells * A * B) return;
  const int b = (int)(row % B);
  const int a = (int)((row / B) % A);
  const int64_t cell = row / ((int64_t)A * B);
  const int64_t gcell = gcell0 + cell;
  const int Lq = net.Lq, L = net.L, E = net.E;
  float u[kMaxLq];
  float s1 = 0.f, s2 = 0.f;
  for (int l0 = 0; l0 < Lq; l0 += 4) {
    float e[4];
    if (noise != nullptr) {
      const float* np_ = noise + (((int64_t)b * A + a) * noise_n + noise_cell0 + cell) * Lq + l0;
#pragma unroll
      for (int j = 0; j < 4; ++j) e[j] = (l0 + j < Lq) ? __ldg(np_ + j) : 0.f;
    } else {
      const float4 g = philox_normal4((uint32_t)gcell, (uint32_t)a, (uint32_t)b,
                                      (uint32_t)(l0 >> 2) | (stream << 16) | ((uint32_t)(gcell >> 32) << 20), seed);
      e[0] = g.x; e[1] = g.y; e[2] = g.z; e[3] = g.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (l0 + j < Lq) {
        const float v = fmaf(__ldg(scale + cell * Lq + l0 + j), e[j], __ldg(mean + cell * Lq + l0 + j));
        u[l0 + j] = v;
        if (u_draw_out != nullptr) u_draw_out[row * Lq + l0 + j] = v;
        s1 += v;
        s2 = fmaf(v, v, s2);
      }
    }
  }
  // z_base from the raw draw
  float* zb = zbase_out + row * L;
  if (net.has_zbase) {
    for (int c = 0; c < L; ++c) {
      float acc = __ldg(net.zbase.b + c);
      for (int l = 0; l < Lq; ++l) acc = fmaf(u[l], __ldg(net.zbase.w + (size_t)l * net.zbase.ld + c), acc);
      zb[c] = acc;
    }
  } else {
    for (int c = 0; c < L; ++c) zb[c] = u[c];
  }
  // LayerNorm(u)
  const float m = s1 / (float)Lq;
  const float var = fmaxf(s2 / (float)Lq - m * m, 0.f);
  const float rstd = 1.0f / sqrtf(var + kLnEps);
  float* ul = u_ln_out + row * Lq;
  for (int l = 0; l < Lq; ++l) {
    u[l] = (u[l] - m) * (rstd * __ldg(net.u_ln.scale + l)) + __ldg(net.u_ln.bias + l);
    ul[l] = u[l];
  }
  float* qt = q_tok_out + row * E;
  for (int e = 0; e < E; ++e) {
    float acc = 0.f;
    for (int l = 0; l < Lq; ++l) acc = fmaf(u[l], __ldg(net.q_tok.w + (size_t)l * net.q_tok.ld + e), acc);
    qt[e] = acc;
  }
}

// use_map=False (reference _module.py:326-329): eps ~ Normal(loc, softplus(raw) + 1e-3) instead of its mean.  z holds
// z_base + loc, raw the pre-softplus scale (CellBuf::raw_out); adds (softplus(raw) + 1e-3) * noise in place.  Row layout
// and noise indexing as k_sample_rows (explicit noise[(b * A + a)][cell][L], or Philox stream `stream`).
__global__ void __launch_bounds__(128)
k_eps_noise(float* __restrict__ z, const float* __restrict__ raw, int64_t n_cells, int A, int B, int L,
            const float* __restrict__ noise, int64_t noise_cell0, int64_t noise_n, int64_t gcell0, uint64_t seed, uint32_t stream) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_cells * A * B) return;
  const int b = (int)(row % B);
  const int a = (int)((row / B) % A);
  const int64_t cell = row / ((int64_t)A * B);
  const int64_t gcell = gcell0 + cell;
  for (int l0 = 0; l0 < L; l0 += 4) {
    float e[4];
    if (noise != nullptr) {
      const float* np_ = noise + (((int64_t)b * A + a) * noise_n + noise_cell0 + cell) * L + l0;
#pragma unroll
      for (int j = 0; j < 4; ++j) e[j] = (l0 + j < L) ? __ldg(np_ + j) : 0.f;
    } else {
      const float4 g = philox_normal4((uint32_t)gcell, (uint32_t)a, (uint32_t)b,
                                      (uint32_t)(l0 >> 2) | (stream << 16) | ((uint32_t)(gcell >> 32) << 20), seed);
      e[0] = g.x; e[1] = g.y; e[2] = g.z; e[3] = g.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (l0 + j < L) {
        const float v = raw[row * L + l0 + j];
        const float sp = (v > 20.f ? v : log1pf(expf(v))) + 1e-3f;
        z[row * L + l0 + j] = fmaf(sp, e[j], z[row * L + l0 + j]);
      }
    }
  }
}

// group code of virtual cell (cell, m): code * mc + m  (sampled_distances keep the mc axis, so a category's sum is
// [mc, S, S], reference _model.py:469-484 applied to a (cell, mc, S, S) DataArray)
__global__ void k_virtual_codes(const int32_t* __restrict__ codes, int64_t n_cells, int mc, int n_groups,
                                int32_t* __restrict__ vcodes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cells * mc) return;
  const int g = codes[i / mc];
  vcodes[i] = (g < 0 || g >= n_groups) ? -1 : g * mc + (int)(i % mc);
}

// local baseline (reference _model.py:535-540): zb [n][2 mc][L] -> mean and population variance over m of
// || zb[m] - zb[m + mc] ||_2.  One warp per cell.
__global__ void k_baseline_stats(const float* __restrict__ zb, int64_t n_cells, int mc, int L, float* __restrict__ mu,
                                 float* __restrict__ var) {
  const int64_t cell = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (cell >= n_cells) return;
  const int lane = threadIdx.x & 31;
  const float* z = zb + cell * 2 * mc * L;
  auto dist_m = [&](int m) {
    float acc = 0.f;
    for (int l = lane; l < L; l += 32) {
      const float d = z[m * L + l] - z[(m + mc) * L + l];
      acc = fmaf(d, d, acc);
    }
    return sqrtf(warp_sum(acc));
  };
  float s1 = 0.f;
  for (int m = 0; m < mc; ++m) s1 += dist_m(m);
  const float mean = s1 / (float)mc;
  float v = 0.f;  // second pass (mc * L values): cancellation-free variance
  for (int m = 0; m < mc; ++m) {
    const float dd = dist_m(m) - mean;
    v = fmaf(dd, dd, v);
  }
  if (lane == 0) mu[cell] = mean;
  if (lane == 0) var[cell] = v / (float)mc;
}

// normalized distances (reference _model.py:447-450): out[cell] = mean_m (D[cell, m] - mu[cell]) / sqrt(var[cell]),
// written per cell and / or summed per group.  Worker = (contiguous range of sorted positions) x (256 entries of
// the S x S block); runs of equal composite code accumulate in a register and flush with float64 atomics.
__global__ void __launch_bounds__(256)
k_normalize_mc(const float* __restrict__ dv, const float* __restrict__ mu, const float* __restrict__ var, int64_t n_cells,
               int mc, int SS, float* __restrict__ cell_out, GroupPlan gp, int cells_per_worker) {
  const int e = blockIdx.y * 256 + threadIdx.x;
  if (e >= SS) return;
  const int64_t p0 = (int64_t)blockIdx.x * cells_per_worker;
  const int64_t p1 = min(n_cells, p0 + cells_per_worker);
  const int S = (int)(sqrtf((float)SS) + 0.5f);
  float acc = 0.f;
  int cur_code = -1;
  int64_t cur_cell = -1;
  const float inv_mc = 1.0f / (float)mc;
  for (int64_t p = p0; p < p1; ++p) {
    const int64_t cell = gp.order ? gp.order[p] : p;
    const int code = gp.n_keys ? gp.comp_sorted[p] : 0;
    if (gp.n_keys && cur_cell >= 0 && code != cur_code) {
      flush_entry(gp, cur_cell, S, e / S, e % S, acc);
      acc = 0.f;
    }
    cur_code = code; cur_cell = cell;
    const float m0 = mu[cell], rs = 1.0f / sqrtf(var[cell]);
    float sum = 0.f;
    for (int m = 0; m < mc; ++m) sum += (dv[((size_t)cell * mc + m) * SS + e] - m0) * rs;
    const float v = sum * inv_mc;
    if (cell_out) cell_out[(size_t)cell * SS + e] = v;
    acc += v;
  }
  if (gp.n_keys && cur_cell >= 0) flush_entry(gp, cur_cell, S, e / S, e % S, acc);
}

// ----------------------------------------------------------------------------------------
// Differential-expression eps stage (reference _model.py:1196-1215, 1283).
// eps [n_cells * mc][S][L] are the residuals of the chain for every (cell, draw).  Per (cell, draw a, latent d):
//   x_s = (eps[s, d] - mean_s eps[., d]) / (1e-6 + std_s eps[., d])      over the KEPT samples (population std)
//   beta[k, d] = sum_s Amat[cell][k, s] x_s                                MLE of the per-cell linear model
//   bn[l, d]   = sum_k beta[k, d] prefactor[cell][k, l]                    whitened coefficients
//   effect_size[cell, l] = mean_a sum_d bn[l, d]^2 ;  beta_out[cell, k, d] = mean_a beta[k, d] * std
// One CTA (4 warps) per cell: warps take draws round-robin, lanes take latent dims; per-warp accumulators in shared
// memory are combined in a fixed order at the end (run-to-run deterministic).
// ----------------------------------------------------------------------------------------
constexpr int kDeWarps = 4;
constexpr int kDeMaxCov = 64;

inline size_t de_stats_smem(int K, int Sk, int L) {
  const int K8 = (K + 7) / 8 * 8, Lp = (L + 31) / 32 * 32;
  return ((size_t)K8 * Sk + (size_t)K * K + (size_t)kDeWarps * K8 * 32 + (size_t)kDeWarps * K * Lp + (size_t)kDeWarps * K) * sizeof(float) +
         (size_t)Sk * sizeof(int);
}

__global__ void __launch_bounds__(kDeWarps * 32)
k_de_stats(const float* __restrict__ eps, int64_t n_cells, int mc, int S, int L, const int32_t* __restrict__ kept, int Sk,
           int K, const float* __restrict__ Amat, const float* __restrict__ pref, int per_cell,
           float* __restrict__ beta_out, float* __restrict__ effect_out) {
  extern __shared__ __align__(16) float de_sm[];
  const int K8 = (K + 7) / 8 * 8, Lp = (L + 31) / 32 * 32;
  float* A = de_sm;                              // [K8][Sk], rows >= K zero
  float* P = A + (size_t)K8 * Sk;                // [K][K]
  float* bw = P + (size_t)K * K;                 // [warps][K8][32] betas of the current (draw, 32 latent dims)
  float* bacc = bw + (size_t)kDeWarps * K8 * 32; // [warps][K][Lp]
  float* tacc = bacc + (size_t)kDeWarps * K * Lp;// [warps][K]
  int* ks = reinterpret_cast<int*>(tacc + kDeWarps * K);
  const int64_t cell = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* Ac = Amat + (per_cell ? (size_t)cell * K * Sk : 0);
  const float* Pc = pref + (per_cell ? (size_t)cell * K * K : 0);
  for (int i = tid; i < K8 * Sk; i += blockDim.x) A[i] = (i < K * Sk) ? __ldg(Ac + i) : 0.f;
  for (int i = tid; i < K * K; i += blockDim.x) P[i] = __ldg(Pc + i);
  for (int i = tid; i < kDeWarps * K * Lp; i += blockDim.x) bacc[i] = 0.f;
  for (int i = tid; i < kDeWarps * K; i += blockDim.x) tacc[i] = 0.f;
  for (int i = tid; i < Sk; i += blockDim.x) ks[i] = kept ? kept[i] : i;
  __syncthreads();
  float* bww = bw + (size_t)warp * K8 * 32;
  float* baw = bacc + (size_t)warp * K * Lp;
  const float inv_sk = 1.0f / (float)Sk;
  for (int a = warp; a < mc; a += kDeWarps) {
    const float* base = eps + ((size_t)(cell * mc + a) * S) * L;
    for (int d0 = 0; d0 < L; d0 += 32) {
      const int d = d0 + lane;
      const bool act = d < L;
      const float* col = base + (act ? d : 0);
      float s1 = 0.f;
      for (int j = 0; j < Sk; ++j) s1 += col[(size_t)ks[j] * L];
      const float mean = s1 * inv_sk;
      float s2 = 0.f;
      for (int j = 0; j < Sk; ++j) { const float dv = col[(size_t)ks[j] * L] - mean; s2 = fmaf(dv, dv, s2); }
      const float sd = sqrtf(s2 * inv_sk);
      const float rs = 1.0f / (1e-6f + sd);
      for (int k0 = 0; k0 < K8; k0 += 8) {
        float b[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < Sk; ++j) {
          const float x = (col[(size_t)ks[j] * L] - mean) * rs;
#pragma unroll
          for (int q = 0; q < 8; ++q) b[q] = fmaf(A[(size_t)(k0 + q) * Sk + j], x, b[q]);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) bww[(k0 + q) * 32 + lane] = b[q];
      }
      __syncwarp();
      if (act)
        for (int k = 0; k < K; ++k) baw[k * Lp + d] = fmaf(bww[k * 32 + lane], sd, baw[k * Lp + d]);
      for (int l = 0; l < K; ++l) {
        float bn = 0.f;
        for (int k = 0; k < K; ++k) bn = fmaf(bww[k * 32 + lane], P[k * K + l], bn);
        const float t = warp_sum(act ? bn * bn : 0.f);
        if (lane == 0) tacc[warp * K + l] += t;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  const float inv_mc = 1.0f / (float)mc;
  for (int i = tid; i < K * L; i += blockDim.x) {
    const int k = i / L, d = i % L;
    float v = 0.f;
    for (int w = 0; w < kDeWarps; ++w) v += bacc[((size_t)w * K + k) * Lp + d];
    beta_out[(size_t)cell * K * L + i] = v * inv_mc;
  }
  for (int l = tid; l < K; l += blockDim.x) {
    float v = 0.f;
    for (int w = 0; w < kDeWarps; ++w) v += tacc[w * K + l];
    effect_out[(size_t)cell * K + l] = v * inv_mc;
  }
}

}  // namespace mrvi
