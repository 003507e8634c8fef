// tcgen05 issue / TMEM / table-gather microbenchmarks on B200 (design input for k_chain_fused_tc, round 2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/mma_microbench tools/mma_microbench.cu
// Questions: (1) what does one tcgen05.mma cost the issuing thread, and how long does a train of n MMAs take to
// complete, as a function of N, SS vs TS operands, and one vs two concurrent issuing warps; (2) tcgen05.ld
// throughput with 8 / 16 warps; (3) 16 independent 16-byte L2 gathers per thread (the attention table) against
// 256 MUFU.EX2 + FMAs per thread (the direct attention evaluation).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../mrvi_b200/csrc/tc_kernels.cuh"

using namespace mrvi;

struct Res { long long issue, done; };

// MODE 0: SS (A, B in shared memory); MODE 1: TS (A in TMEM).  n_mma is rounded to groups of 4 k-steps.
// (elect_one comes from tc_kernels.cuh)

// UNI 1: the issuing branch is warp-uniform (warp index through a shuffle, elect.sync inside) so that descriptors
// live in uniform registers; UNI 0: `if (lane == 0 && ...)` as the round-1 kernels do.
template <int N, int MODE, int UNI>
__global__ void __launch_bounds__(256, 1) k_mma(int n_mma, int n_issuers, int commit_every, Res* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 96 * 1024 + 64);
  const int tid = threadIdx.x;
  const int warp = UNI ? __shfl_sync(0xffffffffu, tid >> 5, 0) : (tid >> 5);
  for (int i = tid; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  fence_proxy_async();
  __syncthreads();
  bool go;
  if (UNI) go = warp < n_issuers; else go = (tid & 31) == 0 && warp < n_issuers;
  if (go && (!UNI || elect_one())) {
    const uint32_t a = smem_u32(smem), b = a + 32 * 1024;  // A: [128 x 128] f16 K-major; B: [N x 64]
    constexpr uint32_t idesc = umma_idesc_f16(N);
    const uint32_t d = tb + warp * 256;
    const uint32_t at = tb + (N == 256 ? 256 : 128) + warp * 256;
    uint32_t phase = 0;
    tc_fence_after();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t bd = umma_desc(b + ks * 2 * (uint32_t)N * 16, (uint32_t)N * 16, 128);
        if (MODE == 0) umma_f16(d, umma_desc(a + ks * 2 * 2048, 2048, 128), bd, idesc, (i | ks) ? 1u : 0u);
        else umma_f16_ts(d, at + ks * 8, bd, idesc, (i | ks) ? 1u : 0u);
      }
      if (commit_every > 0 && (i + 4) % commit_every == 0 && i + 4 < n_mma) {
        umma_commit(&bar[warp]);
        mbar_wait(&bar[warp], phase); phase ^= 1;
        tc_fence_after();
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], phase);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[warp].issue = t1 - t0; out[warp].done = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int N, int MODE, int UNI>
void run_mma(Res* d_res) {
  const int smem = 96 * 1024 + 128;
  cudaFuncSetAttribute(k_mma<N, MODE, UNI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int n : {4, 8, 32, 96})
    for (int iss : {1, 2}) {
      if (iss == 2 && N == 256) continue;
      k_mma<N, MODE, UNI><<<148, 256, smem>>>(n, iss, 0, d_res);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s (N=%d mode=%d n=%d)\n", cudaGetErrorString(e), N, MODE, n); exit(1); }
      Res r[2];
      cudaMemcpy(r, d_res, sizeof(r), cudaMemcpyDeviceToHost);
      printf("%s uni=%d N=%3d n=%3d issuers=%d: issue %6lld clk (%5.1f/mma)  done %6lld clk (%5.1f/mma)\n", MODE ? "TS" : "SS", UNI, N, n, iss,
             r[0].issue, (double)r[0].issue / n, r[0].done, (double)r[0].done / n);
    }
  // trains separated by commit + wait (the chain's structure): 8 trains of 12
  k_mma<N, MODE, UNI><<<148, 256, smem>>>(96, 1, 12, d_res);
  cudaDeviceSynchronize();
  Res r;
  cudaMemcpy(&r, d_res, sizeof(r), cudaMemcpyDeviceToHost);
  printf("%s uni=%d N=%3d 8 trains of 12 MMAs, commit + wait after each: %lld clk = %.0f per train\n", MODE ? "TS" : "SS", UNI, N, r.done, r.done / 8.0);
}

// TMEM read throughput: every warp loads 32 columns x 32 lanes repeatedly
__global__ void __launch_bounds__(512, 1) k_ldtm(int iters, long long* out, float* sink) {
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    float v[32];
    tmem_ld32(t + (i & 1) * 32, v);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 32; q += 8) acc += v[q];
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 1234.5f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}

// attention feature evaluation, direct: per thread 16 features x 16 exponentials (as k_chain_fused_tc)
__global__ void __launch_bounds__(512, 1) k_att_direct(int iters, const float* __restrict__ kvl, long long* out, float* sink) {
  const int tid = threadIdx.x;
  float2 kv2[8];
  for (int q = 0; q < 8; ++q) kv2[q] = make_float2(kvl[(tid & 127) * 16 + 2 * q], kvl[(tid & 127) * 16 + 2 * q + 1]);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll 1
    for (int f = 0; f < 16; ++f) {
      const float t = 0.37f * (float)(f - 8) + 0.001f * (float)it;
      const float2 t2 = make_float2(t, t), off2 = make_float2(-2.f, -2.f);
      float2 den = make_float2(0.f, 0.f), num = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 a = __ffma2_rn(t2, kv2[j], off2);
        const float2 ex = make_float2(ex2_approx(a.x), ex2_approx(a.y));
        den = __fadd2_rn(den, ex);
        num = __ffma2_rn(ex, kv2[j], num);
      }
      acc += (num.x + num.y) * rcp_approx(den.x + den.y);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 1234.5f) sink[0] = acc;
}

// the same features from a per-sample cubic-Hermite table in L2: node k of sample s = float2 (f, h f'); one 16-byte
// gather (two adjacent nodes) + ~10 FMAs per feature
__global__ void __launch_bounds__(512, 1) k_att_table(int iters, const float4* __restrict__ table, int nodes, long long* out, float* sink) {
  const int tid = threadIdx.x;
  const int s = (tid + blockIdx.x * 7) & 127;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float4 nd[16];
    float fr[16];
#pragma unroll
    for (int f = 0; f < 16; ++f) {
      const float t = 0.37f * (float)(f - 8) + 0.013f * (float)((it * 31 + blockIdx.x) & 63);
      const float x = fminf(fmaxf((t + 8.f) * 16.f, 0.f), (float)(nodes - 2));
      const int k = (int)x;
      fr[f] = x - (float)k;
      nd[f] = __ldg(table + (size_t)s * (nodes / 1) + k);
    }
#pragma unroll
    for (int f = 0; f < 16; ++f) {
      const float u = fr[f], f0 = nd[f].x, d0 = nd[f].y, f1 = nd[f].z, d1 = nd[f].w;
      // cubic Hermite: f0 + u (d0 + u (c2 + u c3)),  c2 = 3 df - 2 d0 - d1, c3 = d0 + d1 - 2 df
      const float df = f1 - f0;
      const float c3 = fmaf(-2.f, df, d0 + d1), c2 = fmaf(3.f, df, fmaf(-2.f, d0, -d1));
      acc += fmaf(u, fmaf(u, fmaf(u, c3, c2), d0), f0);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 1234.5f) sink[0] = acc;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
  Res* d_res; long long* d_t; float* d_sink;
  CK(cudaMalloc(&d_res, sizeof(Res) * 8)); CK(cudaMalloc(&d_t, 64)); CK(cudaMalloc(&d_sink, 64));
  printf("== tcgen05.mma trains (one CTA per SM, 148 CTAs; clocks of CTA 0)\n");
  run_mma<32, 0, 1>(d_res); run_mma<64, 0, 1>(d_res); run_mma<128, 0, 1>(d_res); run_mma<256, 0, 1>(d_res);
  run_mma<32, 1, 1>(d_res); run_mma<64, 1, 1>(d_res); run_mma<128, 1, 1>(d_res);
  run_mma<32, 0, 0>(d_res); run_mma<128, 0, 0>(d_res); run_mma<32, 1, 0>(d_res); run_mma<128, 1, 0>(d_res);
  printf("== tcgen05.ld 32x32b.x32 throughput\n");
  for (int threads : {128, 256, 512}) {
    k_ldtm<<<148, threads>>>(2048, d_t, d_sink);
    CK(cudaDeviceSynchronize());
    long long t; CK(cudaMemcpy(&t, d_t, 8, cudaMemcpyDeviceToHost));
    const double bytes = (double)threads / 32 * 2048 * 32 * 32 * 4;
    printf("  %2d warps: %lld clk for %d loads/warp = %.1f clk per load per warp, %.1f B/clk/SM\n", threads / 32, t, 2048, (double)t / 2048, bytes / t);
  }

  printf("== attention features: direct (256 ex2 per thread) vs table gather (16 x 16 B from L2), 512 threads/CTA, 148 CTAs\n");
  std::vector<float> kv(128 * 16);
  for (size_t i = 0; i < kv.size(); ++i) kv[i] = 1.4427f * (float)((int)(i * 2654435761u % 2001) - 1000) / 700.f;
  float* d_kv; CK(cudaMalloc(&d_kv, kv.size() * 4)); CK(cudaMemcpy(d_kv, kv.data(), kv.size() * 4, cudaMemcpyHostToDevice));
  k_att_direct<<<148, 512>>>(64, d_kv, d_t, d_sink);
  CK(cudaDeviceSynchronize());
  { long long t; CK(cudaMemcpy(&t, d_t, 8, cudaMemcpyDeviceToHost)); printf("  direct: %.0f clk per 16-feature batch per CTA (512 threads)\n", t / 64.0); }
  for (int nodes : {256, 512, 1024}) {
    float4* d_tab; CK(cudaMalloc(&d_tab, (size_t)128 * nodes * 16 + 64)); CK(cudaMemset(d_tab, 0, (size_t)128 * nodes * 16 + 64));
    k_att_table<<<148, 512>>>(64, d_tab, nodes, d_t, d_sink);
    CK(cudaDeviceSynchronize());
    k_att_table<<<148, 512>>>(64, d_tab, nodes, d_t, d_sink);
    CK(cudaDeviceSynchronize());
    long long t; CK(cudaMemcpy(&t, d_t, 8, cudaMemcpyDeviceToHost));
    printf("  table %4d nodes/sample (%.0f KB): %.0f clk per 16-feature batch per CTA (512 threads)\n", nodes, 128.0 * nodes * 16 / 1024, t / 64.0);
    CK(cudaFree(d_tab));
  }
  printf("done\n");
  return 0;
}
