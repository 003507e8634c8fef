// Pipe-throughput and approx-accuracy microbenchmarks on B200 (design input for the TC epilogues).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITERS 4096
template <int OP>
__global__ void k(float* out, float seed) {
  float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float b = 1.0001f, c = 0.5f;
  unsigned long long p0, p1, p2, p3;
  if (OP == 1 || OP == 9 || OP == 10) {
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
  }
  unsigned long long pb, pc;
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(pc) : "f"(c));
  __half2 h0 = __floats2half2_rn(a0, a1), h1 = __floats2half2_rn(a2, a3), h2 = __floats2half2_rn(a4, a5), h3 = __floats2half2_rn(a6, a7);
  const __half2 hb = __floats2half2_rn(0.999f, 0.999f), hc = __floats2half2_rn(0.01f, 0.01f);
#pragma unroll 1
  for (int i = 0; i < ITERS; ++i) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (OP == 0) {  // FFMA reg form
        a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
        a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
      } else if (OP == 1) {  // FFMA2
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p0) : "l"(pb), "l"(pc));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p1) : "l"(pb), "l"(pc));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p2) : "l"(pb), "l"(pc));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p3) : "l"(pb), "l"(pc));
      } else if (OP == 2) {  // MUFU.EX2
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
      } else if (OP == 3) {  // MUFU.TANH
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a0)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a1));
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a2)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a3));
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a4)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a5));
        asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a6)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a7));
      } else if (OP == 4) {  // MUFU.RCP
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a1));
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a3));
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a5));
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a7));
      } else if (OP == 5) {  // HFMA2
        h0 = __hfma2(h0, hb, hc); h1 = __hfma2(h1, hb, hc); h2 = __hfma2(h2, hb, hc); h3 = __hfma2(h3, hb, hc);
        h0 = __hfma2(h0, hb, hc); h1 = __hfma2(h1, hb, hc); h2 = __hfma2(h2, hb, hc); h3 = __hfma2(h3, hb, hc);
      } else if (OP == 6) {  // FADD
        a0 += b; a1 += b; a2 += b; a3 += b; a4 += b; a5 += b; a6 += b; a7 += b;
      } else if (OP == 7) {  // FMUL
        a0 *= b; a1 *= b; a2 *= b; a3 *= b; a4 *= b; a5 *= b; a6 *= b; a7 *= b;
      } else if (OP == 8) {  // mixed: 4 FFMA + 4 MUFU.EX2 (dual issue?)
        a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
      } else if (OP == 9) {  // ADD2
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(pb)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pb));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(pb)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pb));
      } else if (OP == 10) {  // MUL2
        asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p0) : "l"(pb)); asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p1) : "l"(pb));
        asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p2) : "l"(pb)); asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p3) : "l"(pb));
      } else if (OP == 11) {  // ex2.approx.f16x2
        unsigned u0 = *(unsigned*)&h0, u1 = *(unsigned*)&h1, u2 = *(unsigned*)&h2, u3 = *(unsigned*)&h3;
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u1));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u2)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u3));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u1));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u2)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u3));
        *(unsigned*)&h0 = u0; *(unsigned*)&h1 = u1; *(unsigned*)&h2 = u2; *(unsigned*)&h3 = u3;
      } else if (OP == 12) {  // cvt.rn.f16x2.f32 (pack)
        unsigned u;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a0), "f"(a1)); a0 += __uint_as_float(u & 0x3f800000u);
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a2), "f"(a3)); a2 += __uint_as_float(u & 0x3f800000u);
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a4), "f"(a5)); a4 += __uint_as_float(u & 0x3f800000u);
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a6), "f"(a7)); a6 += __uint_as_float(u & 0x3f800000u);
      }
    }
  }
  if (OP == 1 || OP == 9 || OP == 10) {
    float x, y;
    asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(p0 ^ p1 ^ p2 ^ p3));
    a0 = x + y;
  }
  float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + __low2float(h0) + __low2float(h1) + __low2float(h2) + __low2float(h3);
  if (r == 12345.678f) out[0] = r;
}

template <int OP>
void run(const char* name, double lane_ops_per_iter_per_thread) {
  float* out; cudaMalloc(&out, 4);
  int blocks = 148 * 8, threads = 256;
  k<OP><<<blocks, threads>>>(out, 0.5f);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<blocks, threads>>>(out, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * ITERS * 4 * lane_ops_per_iter_per_thread;
  double per_clk_sm = ops / (ms * 1e-3) / 148 / 1.965e9;
  printf("%-28s %8.3f ms  %7.1f lane-results/clk/SM (at 1965 MHz)  err=%s\n", name, ms, per_clk_sm, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

__global__ void k_acc(const float* x, float* t_approx, float* e_approx, float* r_approx, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i], t, e, r;
  asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(v));
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v));
  asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + fabsf(v)));
  t_approx[i] = t; e_approx[i] = e; r_approx[i] = r;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s sm_%d%d SMs=%d clock=%d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
  run<0>("FFMA (8 indep)", 8);
  run<1>("FFMA2 fma.rn.f32x2 (4x2)", 8);
  run<6>("FADD", 8);
  run<7>("FMUL", 8);
  run<9>("FADD2 add.rn.f32x2", 8);
  run<10>("FMUL2 mul.rn.f32x2", 8);
  run<2>("MUFU.EX2", 8);
  run<3>("MUFU.TANH", 8);
  run<4>("MUFU.RCP", 8);
  run<11>("MUFU.EX2.f16x2 (per f16)", 16);
  run<5>("HFMA2 (per f16 result)", 16);
  run<8>("4 FFMA + 4 EX2 mixed", 8);
  run<12>("cvt.f16x2.f32 + FADD (pairs)", 4);
  // accuracy
  const int n = 1 << 20;
  float *x, *t, *e, *r; cudaMallocManaged(&x, n * 4); cudaMallocManaged(&t, n * 4); cudaMallocManaged(&e, n * 4); cudaMallocManaged(&r, n * 4);
  for (int i = 0; i < n; ++i) x[i] = -8.0f + 16.0f * (float)i / n;
  k_acc<<<n / 256, 256>>>(x, t, e, r, n); cudaDeviceSynchronize();
  double mt = 0, mta = 0, me = 0, mr = 0, jump = 0;
  for (int i = 0; i < n; ++i) {
    double tr = tanh((double)x[i]);
    double rel = fabs(t[i] - tr) / fmax(fabs(tr), 1e-30);
    if (fabs(x[i]) > 1e-3) mt = fmax(mt, rel);
    mta = fmax(mta, fabs(t[i] - tr));
    me = fmax(me, fabs(e[i] - exp2((double)x[i])) / exp2((double)x[i]));
    mr = fmax(mr, fabs(r[i] - 1.0 / (1.0 + fabs((double)x[i]))) * (1.0 + fabs((double)x[i])));
    if (i > 0) jump = fmax(jump, fabs((t[i] - tanh((double)x[i])) - (t[i - 1] - tanh((double)x[i - 1]))));
  }
  printf("tanh.approx: max rel err %.3e  max abs err %.3e  max neighbour error jump %.3e\n", mt, mta, jump);
  printf("ex2.approx: max rel err %.3e   rcp.approx: max rel err %.3e\n", me, mr);
  return 0;
}
